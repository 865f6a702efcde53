// host_world.cpp — host-side scene construction behind the C ABI (no CUDA in this file).
//
// Mirrors the reference's builder surface:
//   World::new / push_object / push_light                       main.rs:161-178
//   ObjectProxy::push_triangle / push_sphere / push_triangles   main.rs:705-728
//   triangle() / square()  (flat normals)                       main.rs:730-746
//   load_obj()  (tobj: first model, fan triangulation)          main.rs:778-807
//   the scene literal and camera of main()                      main.rs:810-1083
// Built with -ffp-contract=off: every float here rounds exactly like the Rust code.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200rt.h"
#include "hvec.h"

using namespace b200rt_host;

struct b200rt_world {
    std::vector<b200rt_material> materials;  // World.objects (Object = {material}), primitives.rs:8-10
    std::vector<b200rt_triangle> triangles;  // World.triangles, global push order
    std::vector<b200rt_sphere> spheres;      // World.spheres
    std::vector<b200rt_light> lights;        // World.lights
};

extern "C" {

b200rt_world* b200rt_world_new(void) { return new b200rt_world(); }
void b200rt_world_free(b200rt_world* w) { delete w; }

int b200rt_world_push_object(b200rt_world* w, const b200rt_material* material) {
    if (!w || !material) return B200RT_ERR_INVALID;
    w->materials.push_back(*material);
    return (int)w->materials.size() - 1;  // ObjectIndex(self.objects.len() - 1), main.rs:169
}

int b200rt_world_push_triangle(b200rt_world* w, uint32_t object_index, const b200rt_vertex v[3]) {
    if (!w || !v || object_index >= w->materials.size()) return B200RT_ERR_INVALID;
    b200rt_triangle t;
    t.vertices[0] = v[0];
    t.vertices[1] = v[1];
    t.vertices[2] = v[2];
    t.object_index = object_index;
    w->triangles.push_back(t);
    return B200RT_OK;
}

int b200rt_world_push_flat_triangle(b200rt_world* w, uint32_t object_index, const float pos[3][3],
                                    const float uv[3][2]) {
    if (!w || !pos || !uv) return B200RT_ERR_INVALID;
    // main.rs:731-733
    V3 a = v3(pos[1]) - v3(pos[0]);
    V3 b = v3(pos[2]) - v3(pos[1]);
    V3 normal = normalize(cross(a, b));
    b200rt_vertex v[3];
    for (int i = 0; i < 3; ++i) {
        std::memcpy(v[i].position, pos[i], sizeof(float) * 3);
        store(v[i].normal, normal);
        v[i].uv[0] = uv[i][0];
        v[i].uv[1] = uv[i][1];
    }
    return b200rt_world_push_triangle(w, object_index, v);
}

int b200rt_world_push_square(b200rt_world* w, uint32_t object_index, const float pos[4][3],
                             const float uv[4][2]) {
    if (!w || !pos || !uv) return B200RT_ERR_INVALID;
    // main.rs:743-744: (0,1,2) then (0,2,3)
    static const int idx[2][3] = {{0, 1, 2}, {0, 2, 3}};
    for (int t = 0; t < 2; ++t) {
        float p[3][3], q[3][2];
        for (int i = 0; i < 3; ++i) {
            std::memcpy(p[i], pos[idx[t][i]], sizeof(float) * 3);
            std::memcpy(q[i], uv[idx[t][i]], sizeof(float) * 2);
        }
        int rc = b200rt_world_push_flat_triangle(w, object_index, p, q);
        if (rc != B200RT_OK) return rc;
    }
    return B200RT_OK;
}

int b200rt_world_push_sphere(b200rt_world* w, uint32_t object_index, const float center[3], float radius) {
    if (!w || !center || object_index >= w->materials.size()) return B200RT_ERR_INVALID;
    b200rt_sphere s;
    std::memcpy(s.center, center, sizeof(float) * 3);
    s.radius = radius;
    s.object_index = object_index;
    w->spheres.push_back(s);
    return B200RT_OK;
}

int b200rt_world_push_light(b200rt_world* w, const b200rt_light* light) {
    if (!w || !light) return B200RT_ERR_INVALID;
    w->lights.push_back(*light);
    return B200RT_OK;
}

int b200rt_world_scene(const b200rt_world* w, b200rt_scene* out) {
    if (!w || !out) return B200RT_ERR_INVALID;
    out->triangles = w->triangles.data();
    out->n_triangles = (uint32_t)w->triangles.size();
    out->spheres = w->spheres.data();
    out->n_spheres = (uint32_t)w->spheres.size();
    out->materials = w->materials.data();
    out->n_materials = (uint32_t)w->materials.size();
    out->lights = w->lights.data();
    out->n_lights = (uint32_t)w->lights.size();
    return B200RT_OK;
}

}  // extern "C"

// ---- OBJ import -------------------------------------------------------------------------------
// What tobj 0.1.6 does for the reference's call (main.rs:784-790): positions are one global pool,
// a model ends when an `o`/`g` statement follows faces, only models[0] is used, faces of any arity
// are fan-triangulated, `a`, `a/b`, `a//c`, `a/b/c` and negative (relative) indices are accepted.
namespace {

struct ObjMesh {
    std::vector<float> positions;           // xyz...
    std::vector<unsigned> indices;          // triangulated, 0-based, first model only
};

// A Wavefront OBJ file as tobj 0.1.6 (Cargo.toml:14) reads it: `v` / `vt` / `vn` pools shared by the whole file, and one
// MODEL per run of faces between `o` / `g` statements (a new name closes the model that has faces).  Faces are
// triangulated as fans (quads: (0,1,2),(0,2,3)); a corner is `v`, `v/vt`, `v//vn` or `v/vt/vn`, 1-based or negative
// (relative to the pool's size at that line).  The reference takes models[0].mesh.positions / indices and nothing else
// (main.rs:787-790); the extended entry point below also offers the other models, `vt` and `vn`.
struct ObjCorner { int v, vt, vn; };        // 0-based, -1 = absent
struct ObjModel {
    std::string name;
    std::vector<ObjCorner> corners;         // 3 per triangle
};
struct ObjFile {
    std::vector<float> positions, texcoords, normals;   // xyz, uv, xyz
    std::vector<ObjModel> models;
};

bool parse_floats(char* q, int n, std::vector<float>& out) {
    for (int k = 0; k < n; ++k) {
        char* end = nullptr;
        const float val = std::strtof(q, &end);          // correctly rounded, like Rust's str::parse::<f32>
        if (end == q) return false;
        out.push_back(val);
        q = end;
    }
    return true;
}

bool resolve_index(long raw, size_t pool, int& out) {
    const long n = (long)pool;
    const long c = raw < 0 ? n + raw : raw - 1;
    if (c < 0 || c >= n) return false;
    out = (int)c;
    return true;
}

bool parse_obj_file(FILE* f, ObjFile& obj) {
    char line[4096];
    ObjModel cur;
    cur.name = "unnamed_object";                         // tobj's name for faces before any o / g
    auto is_sep = [](char c) { return c == ' ' || c == '\t'; };
    auto is_end = [](char c) { return c == '\0' || c == '\n' || c == '\r'; };
    while (std::fgets(line, sizeof line, f)) {
        char* p = line;
        while (is_sep(*p)) ++p;
        if (p[0] == 'v' && is_sep(p[1])) {
            if (!parse_floats(p + 1, 3, obj.positions)) return false;
        } else if (p[0] == 'v' && p[1] == 't' && is_sep(p[2])) {
            if (!parse_floats(p + 2, 2, obj.texcoords)) return false;
        } else if (p[0] == 'v' && p[1] == 'n' && is_sep(p[2])) {
            if (!parse_floats(p + 2, 3, obj.normals)) return false;
        } else if (p[0] == 'f' && is_sep(p[1])) {
            std::vector<ObjCorner> corner;
            char* q = p + 1;
            for (;;) {
                while (is_sep(*q)) ++q;
                if (is_end(*q) || *q == '#') break;
                ObjCorner c{-1, -1, -1};
                char* end = nullptr;
                const long vi = std::strtol(q, &end, 10);
                if (end == q || !resolve_index(vi, obj.positions.size() / 3, c.v)) return false;
                q = end;
                if (*q == '/') {
                    ++q;
                    if (*q != '/' && !is_sep(*q) && !is_end(*q)) {
                        const long ti = std::strtol(q, &end, 10);
                        if (end == q || !resolve_index(ti, obj.texcoords.size() / 2, c.vt)) return false;
                        q = end;
                    }
                    if (*q == '/') {
                        ++q;
                        if (!is_sep(*q) && !is_end(*q)) {
                            const long ni = std::strtol(q, &end, 10);
                            if (end == q || !resolve_index(ni, obj.normals.size() / 3, c.vn)) return false;
                            q = end;
                        }
                    }
                }
                if (!is_sep(*q) && !is_end(*q)) return false;
                corner.push_back(c);
            }
            if (corner.size() < 3) return false;
            for (size_t i = 1; i + 1 < corner.size(); ++i) {
                cur.corners.push_back(corner[0]);
                cur.corners.push_back(corner[i]);
                cur.corners.push_back(corner[i + 1]);
            }
        } else if ((p[0] == 'g' || p[0] == 'o') && (is_sep(p[1]) || is_end(p[1]))) {
            if (!cur.corners.empty()) { obj.models.push_back(cur); cur.corners.clear(); }
            char* q = p + 1;
            while (is_sep(*q)) ++q;
            char* e = q;
            while (!is_end(*e)) ++e;
            while (e > q && is_sep(e[-1])) --e;
            cur.name.assign(q, e);
        }
    }
    if (!cur.corners.empty()) obj.models.push_back(cur);
    return !obj.models.empty();
}

// the reference's view of the file: models[0], positions through the face's `v` indices
bool parse_obj(FILE* f, ObjMesh& mesh) {
    ObjFile obj;
    if (!parse_obj_file(f, obj)) return false;
    mesh.positions = obj.positions;
    for (const ObjCorner& c : obj.models[0].corners) mesh.indices.push_back((unsigned)c.v);
    return true;
}

// Built-in copy of the reference mesh as data: 20 unit vertices of a regular dodecahedron and its
// 12 pentagons as 36 triangles, in the order dodecahedron.obj lists them (1-based there).
const float kDodecaV[20][3] = {
    {-0.57735f, -0.57735f, 0.57735f}, {0.934172f, 0.356822f, 0.0f},    {0.934172f, -0.356822f, 0.0f},
    {-0.934172f, 0.356822f, 0.0f},    {-0.934172f, -0.356822f, 0.0f},  {0.0f, 0.934172f, 0.356822f},
    {0.0f, 0.934172f, -0.356822f},    {0.356822f, 0.0f, -0.934172f},   {-0.356822f, 0.0f, -0.934172f},
    {0.0f, -0.934172f, -0.356822f},   {0.0f, -0.934172f, 0.356822f},   {0.356822f, 0.0f, 0.934172f},
    {-0.356822f, 0.0f, 0.934172f},    {0.57735f, 0.57735f, -0.57735f}, {0.57735f, 0.57735f, 0.57735f},
    {-0.57735f, 0.57735f, -0.57735f}, {-0.57735f, 0.57735f, 0.57735f}, {0.57735f, -0.57735f, -0.57735f},
    {0.57735f, -0.57735f, 0.57735f},  {-0.57735f, -0.57735f, -0.57735f}};
const unsigned char kDodecaF[36][3] = {
    {19, 3, 2},  {12, 19, 2}, {15, 12, 2}, {8, 14, 2},  {18, 8, 2},  {3, 18, 2},  {20, 5, 4},  {9, 20, 4},
    {16, 9, 4},  {13, 17, 4}, {1, 13, 4},  {5, 1, 4},   {7, 16, 4},  {6, 7, 4},   {17, 6, 4},  {6, 15, 2},
    {7, 6, 2},   {14, 7, 2},  {10, 18, 3}, {11, 10, 3}, {19, 11, 3}, {11, 1, 5},  {10, 11, 5}, {20, 10, 5},
    {20, 9, 8},  {10, 20, 8}, {18, 10, 8}, {9, 16, 7},  {8, 9, 7},   {14, 8, 7},  {12, 15, 6}, {13, 12, 6},
    {17, 13, 6}, {13, 1, 11}, {12, 13, 11}, {19, 12, 11}};

int push_mesh(b200rt_world* w, uint32_t object_index, const ObjMesh& mesh, float scale_div, const float offset[3]) {
    const float zero_uv[3][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}};  // main.rs:797-799
    const V3 off = v3(offset);
    int n = 0;
    for (size_t f = 0; f + 2 < mesh.indices.size(); f += 3) {
        float pos[3][3];
        for (int k = 0; k < 3; ++k) {
            const float* src = &mesh.positions[3 * (size_t)mesh.indices[f + k]];
            // main.rs:802: p.position / 3.0 + Vector3::new(0.7, 1.0, -0.5)
            V3 p = v3(src) / scale_div + off;
            store(pos[k], p);
        }
        int rc = b200rt_world_push_flat_triangle(w, object_index, pos, zero_uv);
        if (rc != B200RT_OK) return rc;
        ++n;
    }
    return n;
}

b200rt_material color_material(float dr, float dg, float db, float shiness, float sr, float sg, float sb,
                               float smoothness, float refraction_index, float opaque_decay, float transparency) {
    b200rt_material m;
    std::memset(&m, 0, sizeof m);
    m.kind = B200RT_MATERIAL_COLOR;
    m.normal[0] = 0.0f; m.normal[1] = 0.0f; m.normal[2] = 1.0f;
    m.diffuse_color[0] = dr; m.diffuse_color[1] = dg; m.diffuse_color[2] = db;
    m.shiness = shiness;
    m.specular_color[0] = sr; m.specular_color[1] = sg; m.specular_color[2] = sb;
    m.smoothness = smoothness;
    m.refraction_index = refraction_index;
    m.opaque_decay = opaque_decay;
    m.transparency = transparency;
    m.diffuse_fn = B200RT_DIFFUSE_CONST;
    m.normal_fn = B200RT_NORMAL_CONST;
    return m;
}

struct Quad {
    float pos[4][3];
    float uv[4][2];
};

}  // namespace

extern "C" {

int b200rt_world_load_obj(b200rt_world* w, uint32_t object_index, const char* path, float scale_div,
                          const float offset[3]) {
    if (!w || !path || !offset || object_index >= w->materials.size()) return B200RT_ERR_INVALID;
    FILE* f = std::fopen(path, "r");
    if (!f) return B200RT_ERR_IO;
    ObjMesh mesh;
    bool ok = parse_obj(f, mesh);
    std::fclose(f);
    if (!ok) return B200RT_ERR_IO;
    return push_mesh(w, object_index, mesh, scale_div, offset);
}

// The whole file behind the same transform: any model (or all of them, in file order), and - on request - the file's own
// texture coordinates (`vt`) and vertex normals (`vn`) instead of uv = (0,0) and the flat normal of triangle()
// (main.rs:730-739).  A corner without `vt` / `vn` keeps the reference's value.  Normals are not transformed: the
// reference's placement is a uniform scale and a translation (main.rs:802).
int b200rt_world_load_obj_ex(b200rt_world* w, uint32_t object_index, const char* path, float scale_div,
                             const float offset[3], int32_t model_index, uint32_t flags) {
    if (!w || !path || !offset || object_index >= w->materials.size()) return B200RT_ERR_INVALID;
    if (flags & ~(uint32_t)(B200RT_OBJ_USE_TEXCOORDS | B200RT_OBJ_USE_NORMALS)) return B200RT_ERR_INVALID;
    FILE* f = std::fopen(path, "r");
    if (!f) return B200RT_ERR_IO;
    ObjFile obj;
    const bool ok = parse_obj_file(f, obj);
    std::fclose(f);
    if (!ok) return B200RT_ERR_IO;
    if (model_index >= (int32_t)obj.models.size() || model_index < -1) return B200RT_ERR_INVALID;
    const V3 off = v3(offset);
    int n = 0;
    for (size_t m = 0; m < obj.models.size(); ++m) {
        if (model_index >= 0 && (size_t)model_index != m) continue;
        const std::vector<ObjCorner>& cs = obj.models[m].corners;
        for (size_t t = 0; t + 2 < cs.size(); t += 3) {
            float pos[3][3], uv[3][2];
            bool own_normals = (flags & B200RT_OBJ_USE_NORMALS) != 0u;
            for (int k = 0; k < 3; ++k) {
                const ObjCorner& c = cs[t + k];
                store(pos[k], v3(&obj.positions[3 * (size_t)c.v]) / scale_div + off);          // main.rs:802
                const bool has_uv = (flags & B200RT_OBJ_USE_TEXCOORDS) && c.vt >= 0;
                uv[k][0] = has_uv ? obj.texcoords[2 * (size_t)c.vt] : 0.0f;                     // main.rs:797-799
                uv[k][1] = has_uv ? obj.texcoords[2 * (size_t)c.vt + 1] : 0.0f;
                if (c.vn < 0) own_normals = false;
            }
            int rc;
            if (own_normals) {
                b200rt_vertex v[3];
                for (int k = 0; k < 3; ++k) {
                    std::memcpy(v[k].position, pos[k], sizeof(float) * 3);
                    std::memcpy(v[k].normal, &obj.normals[3 * (size_t)cs[t + k].vn], sizeof(float) * 3);
                    v[k].uv[0] = uv[k][0]; v[k].uv[1] = uv[k][1];
                }
                rc = b200rt_world_push_triangle(w, object_index, v);
            } else {
                rc = b200rt_world_push_flat_triangle(w, object_index, pos, uv);
            }
            if (rc != B200RT_OK) return rc;
            ++n;
        }
    }
    return n;
}

// number of models tobj::load_obj would return for the file (runs of faces between o / g statements), or < 0
int b200rt_obj_model_count(const char* path) {
    if (!path) return B200RT_ERR_INVALID;
    FILE* f = std::fopen(path, "r");
    if (!f) return B200RT_ERR_IO;
    ObjFile obj;
    const bool ok = parse_obj_file(f, obj);
    std::fclose(f);
    return ok ? (int)obj.models.size() : B200RT_ERR_IO;
}

void b200rt_fixture_camera(b200rt_camera* cam) {
    if (!cam) return;
    // main.rs:1077-1083; Deg -> Rad is deg * (PI/180 computed in f64, cast to f32) in cgmath 0.16
    cam->fovy = 60.0f * (float)(3.14159265358979323846 / 180.0);
    cam->center[0] = 2.0f; cam->center[1] = 2.5f; cam->center[2] = 2.0f;
    store(cam->toward, normalize(v3(-1.0f, -1.0f, -1.0f)));
    store(cam->up, normalize(v3(0.0f, 1.0f, 0.0f)));
    cam->near = -0.1f;
}

void b200rt_default_params(b200rt_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->width = 1280;               // main.rs:1084
    p->height = 960;               // main.rs:1085
    p->depth = 5;                  // main.rs:1098
    p->threshold = 0.001f;         // main.rs:467
    p->refract_max_distance = 100.0f;  // main.rs:505
    p->tir_retries = 10;           // main.rs:378
    p->focus = 3.0f;               // main.rs:1147
    p->blur = 0.04f;               // main.rs:1148
    p->seed = 0;
    p->cast_mode = B200RT_CAST_TWO_PHASE;
}

int b200rt_world_fixture(b200rt_world* w, const char* obj_path) {
    if (!w) return B200RT_ERR_INVALID;
    const float obj_offset[3] = {0.7f, 1.0f, -0.5f};
    int rc;

    // object 0: dodecahedron, main.rs:812-825
    b200rt_material m0 = color_material(1.0f, 1.0f, 1.0f, 0.1f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.0f, 0.0f);
    int o0 = b200rt_world_push_object(w, &m0);
    if (obj_path) {
        rc = b200rt_world_load_obj(w, (uint32_t)o0, obj_path, 3.0f, obj_offset);
    } else {
        ObjMesh mesh;
        for (int i = 0; i < 20; ++i)
            for (int k = 0; k < 3; ++k) mesh.positions.push_back(kDodecaV[i][k]);
        for (int f = 0; f < 36; ++f)
            for (int k = 0; k < 3; ++k) mesh.indices.push_back((unsigned)kDodecaF[f][k] - 1u);
        rc = push_mesh(w, (uint32_t)o0, mesh, 3.0f, obj_offset);
    }
    if (rc < 0) return rc;

    // object 1: floor, main.rs:826-844 (uv of corner 3 really is (0,1) again in the reference)
    b200rt_material m1 = color_material(1.0f, 0.8f, 0.6f, 0.5f, 1.0f, 1.0f, 1.0f, 0.01f, 1.0f, 0.0f, 0.0f);
    int o1 = b200rt_world_push_object(w, &m1);
    const Quad floor_q = {{{-2.0f, 0.0f, -2.0f}, {-2.0f, 0.0f, 2.0f}, {2.0f, 0.0f, 2.0f}, {2.0f, 0.0f, -2.0f}},
                          {{0.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}}};
    if ((rc = b200rt_world_push_square(w, (uint32_t)o1, floor_q.pos, floor_q.uv)) < 0) return rc;

    // object 2: striped, bump-mapped wall, main.rs:845-877
    b200rt_material m2 = color_material(0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 0.00001f, 1.0f, 0.0f, 0.0f);
    m2.kind = B200RT_MATERIAL_GENERATIVE;
    m2.diffuse_fn = B200RT_DIFFUSE_STRIPE_V;
    m2.normal_fn = B200RT_NORMAL_SINCOS_U;
    m2.fn_params[0] = 20.0f;
    m2.fn_params[1] = 1.0f; m2.fn_params[2] = 1.0f; m2.fn_params[3] = 1.0f;
    m2.fn_params[4] = 0.5f; m2.fn_params[5] = 0.5f; m2.fn_params[6] = 1.0f;
    m2.fn_params[7] = 10.0f;
    int o2 = b200rt_world_push_object(w, &m2);
    const Quad wall_q = {{{-2.0f, 2.0f, -2.0f}, {-2.0f, 2.0f, 2.0f}, {-2.0f, -2.0f, 2.0f}, {-2.0f, -2.0f, -2.0f}},
                         {{0.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}}};
    if ((rc = b200rt_world_push_square(w, (uint32_t)o2, wall_q.pos, wall_q.uv)) < 0) return rc;

    // objects 3 and 4: two glass slabs, main.rs:879-977.  Same six-quad box pattern:
    //   front z=z1, back z=z0, top y=1.5, bottom y=1.0, left x=-hx, right x=+hx.
    const float uvA[4][2] = {{0.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}};  // first quad of each slab
    const float uvB[4][2] = {{0.0f, 1.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}, {0.0f, 0.0f}};  // the other five
    const float slab[2][3] = {{0.5f, 0.6f, 0.7f}, {0.3f, 0.71f, 0.81f}};               // hx, z0, z1
    for (int s = 0; s < 2; ++s) {
        b200rt_material mg = color_material(1.0f, 0.8f, 0.6f, 1.0f, 1.0f, 1.0f, 1.0f, 0.00001f, 1.6f, 0.1f, 1.0f);
        int og = b200rt_world_push_object(w, &mg);
        const float hx = slab[s][0], z0 = slab[s][1], z1 = slab[s][2];
        const float y0 = 1.0f, y1 = 1.5f;
        const float q0[4][3] = {{hx, y1, z1}, {-hx, y1, z1}, {-hx, y0, z1}, {hx, y0, z1}};     // main.rs:892-897 / 942-947
        const float q1[4][3] = {{hx, y0, z0}, {-hx, y0, z0}, {-hx, y1, z0}, {hx, y1, z0}};     // main.rs:898-903 / 948-953
        const float q2[4][3] = {{hx, y1, z0}, {-hx, y1, z0}, {-hx, y1, z1}, {hx, y1, z1}};     // main.rs:904-909 / 954-959
        const float q3a[4][3] = {{hx, y0, z1}, {-hx, y0, z1}, {-hx, y0, z0}, {hx, y0, z0}};    // bottom
        const float q4[4][3] = {{-hx, y1, z0}, {-hx, y0, z0}, {-hx, y0, z1}, {-hx, y1, z1}};   // left
        const float q5[4][3] = {{hx, y0, z0}, {hx, y1, z0}, {hx, y1, z1}, {hx, y0, z1}};       // right
        if ((rc = b200rt_world_push_square(w, (uint32_t)og, q0, uvA)) < 0) return rc;
        if ((rc = b200rt_world_push_square(w, (uint32_t)og, q1, uvB)) < 0) return rc;
        if ((rc = b200rt_world_push_square(w, (uint32_t)og, q2, uvB)) < 0) return rc;
        if (s == 0) {
            // slab 1 order: bottom (910-915), left (916-921), right (922-927)
            if ((rc = b200rt_world_push_square(w, (uint32_t)og, q3a, uvB)) < 0) return rc;
            if ((rc = b200rt_world_push_square(w, (uint32_t)og, q4, uvB)) < 0) return rc;
        } else {
            // slab 2 order: left (960-965), bottom (966-971), right (972-977)
            if ((rc = b200rt_world_push_square(w, (uint32_t)og, q4, uvB)) < 0) return rc;
            if ((rc = b200rt_world_push_square(w, (uint32_t)og, q3a, uvB)) < 0) return rc;
        }
        if ((rc = b200rt_world_push_square(w, (uint32_t)og, q5, uvB)) < 0) return rc;
    }

    // four spheres, main.rs:979-1056
    const float inv_sqrt3 = 0.5f / std::sqrt(3.0f);
    b200rt_material m5 = color_material(1.0f, 0.2f, 0.2f, 0.2f, 1.0f, 1.0f, 0.0f, 0.2f, 1.0f, 0.0f, 0.0f);
    int o5 = b200rt_world_push_object(w, &m5);
    const float c5[3] = {-0.5f, 0.5f, inv_sqrt3};
    if ((rc = b200rt_world_push_sphere(w, (uint32_t)o5, c5, 0.5f)) < 0) return rc;

    b200rt_material m6 = color_material(1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.001f, 1.12f, 0.3f, 0.96f);
    int o6 = b200rt_world_push_object(w, &m6);
    const float c6[3] = {0.5f, 0.5f, inv_sqrt3};
    if ((rc = b200rt_world_push_sphere(w, (uint32_t)o6, c6, 0.5f)) < 0) return rc;

    b200rt_material m7 = color_material(0.0f, 0.0f, 0.0f, 0.3f, 0.0f, 0.0f, 1.0f, 0.7f, 1.0f, 0.0f, 0.0f);
    m7.kind = B200RT_MATERIAL_GENERATIVE;
    m7.diffuse_fn = B200RT_DIFFUSE_CHECKER_UPV;
    m7.normal_fn = B200RT_NORMAL_CONST;
    m7.fn_params[0] = 10.0f;
    m7.fn_params[1] = 1.0f; m7.fn_params[2] = 0.1f; m7.fn_params[3] = 0.1f;
    m7.fn_params[4] = 0.1f; m7.fn_params[5] = 0.1f; m7.fn_params[6] = 1.0f;
    int o7 = b200rt_world_push_object(w, &m7);
    const float c7[3] = {0.0f, 0.5f, -1.0f / std::sqrt(3.0f)};
    if ((rc = b200rt_world_push_sphere(w, (uint32_t)o7, c7, 0.5f)) < 0) return rc;

    b200rt_material m8 = color_material(0.5f, 1.0f, 0.2f, 0.5f, 1.0f, 1.0f, 1.0f, 0.01f, 1.0f, 0.0f, 0.0f);
    int o8 = b200rt_world_push_object(w, &m8);
    const float c8[3] = {0.0f, 0.5f + std::sqrt(2.0f / 3.0f), 0.0f};
    if ((rc = b200rt_world_push_sphere(w, (uint32_t)o8, c8, 0.5f)) < 0) return rc;

    // three lights, main.rs:1058-1075
    b200rt_light l0;
    std::memset(&l0, 0, sizeof l0);
    l0.kind = B200RT_LIGHT_DIRECTIONAL;
    l0.has_origin = 0;
    store(l0.direction, normalize(v3(-1.0f, -1.0f, 0.0f)));
    l0.color[0] = 1.0f; l0.color[1] = 0.98f; l0.color[2] = 0.95f;
    b200rt_world_push_light(w, &l0);

    b200rt_light l1;
    std::memset(&l1, 0, sizeof l1);
    l1.kind = B200RT_LIGHT_SPOT;
    l1.has_origin = 1;
    l1.origin[0] = 0.0f; l1.origin[1] = 10.0f; l1.origin[2] = 0.0f;
    store(l1.direction, normalize(v3(0.0f, -1.0f, -0.0f)));
    l1.angle = 60.0f * (float)(3.14159265358979323846 / 180.0);
    l1.softness = 1.0f;
    l1.color[0] = 1.0f * 1.0f; l1.color[1] = 0.5f * 1.0f; l1.color[2] = 0.9f * 1.0f;
    b200rt_world_push_light(w, &l1);

    b200rt_light l2;
    std::memset(&l2, 0, sizeof l2);
    l2.kind = B200RT_LIGHT_POINT;
    l2.has_origin = 1;
    l2.origin[0] = 0.0f; l2.origin[1] = 0.1f; l2.origin[2] = 0.0f;
    l2.color[0] = 0.8f; l2.color[1] = 0.8f; l2.color[2] = 1.0f;
    b200rt_world_push_light(w, &l2);
    return B200RT_OK;
}

}  // extern "C"

// ---- write_to_file (main.rs:764-776): RGB8 PNG written to a temporary next to the target, then renamed over it ----
// (the reference writes ./tmp.png and renames: a reader of `name` never sees a half-written image).  The png crate's
// deflate is not reproduced: the IDAT stream uses stored (uncompressed) deflate blocks, which every decoder reads;
// parity is on the decoded pixels.
namespace {

uint32_t crc32_update(uint32_t crc, const unsigned char* data, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ data[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}

bool write_chunk(FILE* f, const char type[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xffffffffu, head.data() + 4, 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<unsigned char> tail;
    put_be32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, 4, f) == 4;
}

}  // namespace

extern "C" int b200rt_write_png_rgb8(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb || width == 0 || height == 0) return B200RT_ERR_INVALID;
    // raw scanlines: filter byte 0 + 3 * width samples
    const size_t stride = 1 + 3 * (size_t)width;
    std::vector<unsigned char> raw(stride * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw[y * stride] = 0;
        std::memcpy(&raw[y * stride + 1], rgb + (size_t)y * 3 * width, 3 * (size_t)width);
    }
    // zlib stream of stored blocks
    std::vector<unsigned char> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;   // Adler-32
    size_t at = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - at);
        z.push_back(at + n == raw.size() ? 1 : 0);
        z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + at, raw.begin() + at + n);
        for (size_t i = at; i < at + n; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        at += n;
    } while (at < raw.size());
    put_be32(z, (b << 16) | a);

    const std::string target(path);
    const size_t slash = target.find_last_of('/');
    const std::string tmp = (slash == std::string::npos ? std::string("./") : target.substr(0, slash + 1)) + "tmp.png";   // main.rs:767
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return B200RT_ERR_IO;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGB (main.rs:770)
    bool ok = std::fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr) && write_chunk(f, "IDAT", z) &&
              write_chunk(f, "IEND", std::vector<unsigned char>());
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { std::remove(tmp.c_str()); return B200RT_ERR_IO; }
    if (std::rename(tmp.c_str(), target.c_str()) != 0) { std::remove(tmp.c_str()); return B200RT_ERR_IO; }   // main.rs:774-775
    return B200RT_OK;
}
