"""The reference's main() (main.rs:809-1173) end to end on the GPU through the C ABI: fixture scene, Whitted frame,
then E epochs of the stochastic tracer added into the progressively renormalised image; writes PATH after every
frame like the reference overwrites ./out.png.

    python tools/render_main.py [WxH] [epochs] [out.png]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

b = g.load_package()
w, h = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1280x960").split("x"))
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 100          # main.rs:1129
path = sys.argv[3] if len(sys.argv) > 3 else "./out.png"
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
t0 = time.perf_counter()
b.render_main(ctx, b.fixture_camera(), b.default_params(width=w, height=h), epochs, out_path=path,
              on_frame=lambda k, u8: print(f"frame {k}: {time.perf_counter() - t0:.3f} s", flush=True) if k in (0, epochs) else None)
print(f"{w}x{h}, {epochs} epochs -> {path}")
