"""The Rust `-sys` declarations (rust/b200rt-sys/src/lib.rs) cannot be compiled in this image (no rustc), so their
#[repr(C)] structs are linted against the ctypes mirror of include/b200rt.h: same structs, same field names in the
same order, same scalar types and array lengths; and every function the crate declares is exported by the library
with the same number of arguments as the ctypes binding declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "rust", "b200rt-sys", "src", "lib.rs")).read()

SCALARS = {"f32": C.c_float, "u32": C.c_uint32, "i32": C.c_int32, "u64": C.c_uint64, "u8": C.c_uint8, "usize": C.c_size_t}


def rust_structs():
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{]*?pub struct (\w+) \{(.*?)\n\}", SRC, re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", m.group(2))
        out[m.group(1)] = [(n, t.strip()) for n, t in fields]
    return out


def matches(rust_type, ctype, pods):
    m = re.fullmatch(r"\[(\w+); (\d+)\]", rust_type)
    if m:
        return (issubclass(ctype, C.Array) and ctype._length_ == int(m.group(2)) and matches(m.group(1), ctype._type_, pods))
    if rust_type.startswith("*const ") or rust_type.startswith("*mut "):
        return issubclass(ctype, (C._Pointer, C.c_void_p)) or ctype in (C.c_void_p, C.c_char_p)
    if rust_type in SCALARS:
        return ctype is SCALARS[rust_type]
    return rust_type in pods and ctype is pods[rust_type]


def test_structs_mirror_the_c_header(b200rt):
    pods = {"b200rt_vertex": b200rt.Vertex, "b200rt_triangle": b200rt.Triangle, "b200rt_sphere": b200rt.Sphere,
            "b200rt_material": b200rt.Material, "b200rt_light": b200rt.Light, "b200rt_scene": b200rt.Scene,
            "b200rt_camera": b200rt.Camera, "b200rt_ray": b200rt.Ray, "b200rt_hit": b200rt.Hit,
            "b200rt_params": b200rt.Params, "b200rt_stats": b200rt.Stats}
    rs = rust_structs()
    assert set(rs) == set(pods)
    for name, cls in pods.items():
        cf = list(cls._fields_)
        assert [n for n, _ in rs[name]] == [n for n, _ in cf], name
        for (rn, rt), (cn, ct) in zip(rs[name], cf):
            assert matches(rt, ct, pods), (name, rn, rt, ct)


def test_declared_functions_are_exported_with_matching_arity(b200rt):
    lib = b200rt.load_library()
    block = re.search(r'extern "C" \{(.*?)\n\}', SRC, re.S).group(1)
    fns = re.findall(r"pub fn (\w+)\((.*?)\)(?: -> [^;]+)?;", block, re.S)
    assert len(fns) >= 30
    for name, args in fns:
        assert name in b200rt.EXPORTED_SYMBOLS, name
        fn = getattr(lib, name)                      # raises if the symbol is missing
        n_args = len([a for a in args.split(",") if a.strip()])
        assert fn.argtypes is not None and len(fn.argtypes) == n_args, (name, n_args, fn.argtypes)


def test_constants_match(b200rt):
    consts = dict(re.findall(r"pub const B200RT_(\w+): (?:c_int|u32) = (-?\d+);", SRC))
    for k, v in consts.items():
        assert getattr(b200rt, k) == int(v), k
