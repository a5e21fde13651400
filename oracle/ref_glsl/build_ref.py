#!/usr/bin/env python
"""Builds oracle/_ref/libddgi_ref.so: the reference's OWN compute shaders
(/root/reference/assets/shaders/probe_pass.comp, compute_pass.comp and everything they
#include), transpiled where they lie into C++ and compiled with g++ against glsl_shim.h.

TEST INFRASTRUCTURE.  Nothing of the reference is copied into the repository: the generated
C++ lives only under oracle/_ref/ (git-ignored); the compiled .so travels to the GPU box.
Used to pin the oracle restatement (oracle/ddgi_oracle.c) against the reference's own text
and to generate tests/golden/*.npz (tests/golden/make_golden.py).

The transpilation is purely lexical — GLSL and C++ share their expression and statement
grammar — and touches only what differs:
  * #include resolution, #version / #extension removal;
  * `layout(...) uniform Block {..} name;` -> `struct Block {..} name;`, image and buffer
    bindings -> shim objects;
  * `out T x` / `inout T x` parameters -> `T& x`;
  * unsuffixed floating literals -> float literals (`0.5` -> `0.5f`: GLSL literals are fp32);
  * `.xyz`-style swizzles -> member calls, `int(x)` / `uint(x)` -> the pinned conversions;
  * `main` -> `shader_main`; file-scope variables with initialisers are re-initialised before
    every invocation (GLSL evaluates them per invocation, probe_pass.comp:55-57);
  * `glsl_count_lookup();` is inserted at the top of getBlockAt so the harness can report the
    reference's own voxel-lookup count per invocation.
The HOST side of the path is pinned the same way: generate_samples and RVPT::generate_probe_rays
(src/rvpt/rvpt.cpp:1145-1224) and struct ProbeRay (src/rvpt/probe.h, included where it lies) are
compiled verbatim against a small glm stand-in (glm_shim/glm/glm.hpp; glm itself is fetched from
the network by the reference's CMake) -> ref_generate_probe_rays.

Further builds restore lines the reference itself has commented out — comment markers removed,
nothing else (RESTORE below):
  ref_probe_pass_hysteresis   the hysteresis blend, probe_pass.comp:298-299;
  ref_probe_pass_lights / ref_compute_pass_lights   the `update_lights();` call at the top of
                              both main()s, probe_pass.comp:254 / compute_pass.comp:174;
  ref_compute_pass_chebyshev  `weight *= chebyshevWeight;`, intersection.glsl:1382.

    python oracle/ref_glsl/build_ref.py [--reference /root/reference] [--keep-going]
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")


def split_comments(src: str):
    """Yields (is_code, text) segments; comments and string literals are left untouched."""
    pat = re.compile(r"//[^\n]*|/\*.*?\*/", re.S)
    pos = 0
    for m in pat.finditer(src):
        if m.start() > pos:
            yield True, src[pos:m.start()]
        yield False, m.group(0)
        pos = m.end()
    if pos < len(src):
        yield True, src[pos:]


def resolve_includes(path: str, shader_dir: str, seen: list[str]) -> str:
    out = []
    for line in open(path, encoding="utf-8", errors="replace").read().split("\n"):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            inc = os.path.join(shader_dir, m.group(1))
            seen.append(m.group(1))
            out.append(f"// ---- begin {m.group(1)}")
            out.append(resolve_includes(inc, shader_dir, seen))
            out.append(f"// ---- end {m.group(1)}")
        elif re.match(r"\s*#(version|extension)\b", line):
            out.append("")
        else:
            out.append(line)
    return "\n".join(out)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
SWIZZLE = re.compile(r"(?<=[\w\)\]])\.([xyzw]{2,4}|[rgba]{2,4})\b(?!\s*\()")


def transpile_code(code: str) -> str:
    # layout qualifiers
    code = re.sub(r"layout\s*\([^)]*\)\s*in\s*;", "", code)
    code = re.sub(r"layout\s*\([^)]*\)\s*buffer\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", r"\1* \2;", code)
    code = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(?:readonly\s+|writeonly\s+)*image2D\s+(\w+)\s*;", r"image2D \1;", code)
    code = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s*\{", r"struct \1 {", code)
    # parameter qualifiers
    code = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", code)
    code = re.sub(r"\bconst\s+in\s+(\w+)\s+(\w+)", r"const \1 \2", code)
    code = re.sub(r"\bin\s+(\w+)\s+(\w+)(?=\s*(?:[,)]|$))", r"\1 \2", code)
    # literals, swizzles, scalar constructors
    code = FLOAT_LIT.sub(lambda m: m.group(1) + "f", code)
    code = SWIZZLE.sub(lambda m: "." + m.group(1) + "()", code)
    code = re.sub(r"\bint\s*\(", "glsl_int(", code)
    code = re.sub(r"\buint\s*\(", "glsl_uint(", code)
    # array sizes taken from a const int table must be constant expressions in C++
    code = re.sub(r"\bconst\s+int\s+(\w+)\s*\[", r"constexpr int \1[", code)
    code = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", code)
    return code


def transpile(src: str) -> str:
    return "".join(transpile_code(t) if is_code else t for is_code, t in split_comments(src))


def file_scope_initialisers(cpp: str) -> list[tuple[str, str]]:
    """(name, expression) of every file-scope `T name = expr;` whose type is a scalar or a
    shim vector: GLSL re-evaluates them for every invocation."""
    out = []
    depth = 0
    code_only = "".join(t if is_code else re.sub(r"[^\n]", " ", t) for is_code, t in split_comments(cpp))
    stmt_start = 0
    i = 0
    n = len(code_only)
    while i < n:
        ch = code_only[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                stmt_start = i + 1
        elif ch == ";" and depth == 0:
            stmt = code_only[stmt_start:i].strip()
            m = re.match(r"^(?:uint|int|float|bool|ivec2|ivec3|vec2|vec3|vec4)\s+(\w+)\s*=\s*(.+)$", stmt, re.S)
            if m and "#" not in stmt:
                out.append((m.group(1), " ".join(m.group(2).split())))
            stmt_start = i + 1
        elif ch == "#":  # preprocessor line: skip to end of line
            j = code_only.find("\n", i)
            i = n if j < 0 else j
            stmt_start = i + 1
        i += 1
    return out


HARNESS_COMMON = r'''
struct RefSettings { int screen_width, screen_height, max_bounces, camera_mode, render_mode, scene; float time; int visualize_probes; };
struct RefField { int probe_count[3]; int side_length; float hysteresis; int sqrt_rays_per_probe; float field_origin[3]; };
'''

HARNESS_PROBE = r'''
namespace %(ns)s {
static void reinit_globals() { %(reinit)s }
static void apply(const RefSettings* s, const RefField* f)
{
    render_settings.screen_width = s->screen_width; render_settings.screen_height = s->screen_height;
    render_settings.max_bounces = s->max_bounces; render_settings.camera_mode = s->camera_mode;
    render_settings.render_mode = s->render_mode; render_settings.scene = s->scene;
    render_settings.time = s->time; render_settings.visualize_probes = s->visualize_probes != 0;
    irradiance_field.probe_count = ivec3(f->probe_count[0], f->probe_count[1], f->probe_count[2]);
    irradiance_field.side_length = f->side_length; irradiance_field.hysteresis = f->hysteresis;
    irradiance_field.sqrt_rays_per_probe = f->sqrt_rays_per_probe;
    irradiance_field.field_origin = vec3(f->field_origin[0], f->field_origin[1], f->field_origin[2]);
}
}  // namespace %(ns)s

// One probe_pass.comp invocation per texel of the W x H probe texture (W, H multiples of the
// tile; invocations of the rounded-up dispatch beyond W or H are not executed, oracle PIN 6).
extern "C" __attribute__((visibility("default")))
void %(entry)s(const RefSettings* s, const RefField* f, const float* rays12, uint32_t n_rays, int W, int H,
                    uint32_t* albedo, uint32_t* distances, float* albedo_f32, uint32_t* lookups)
{
    using namespace %(ns)s;
    apply(s, f);
    std::vector<ProbeRay> list(n_rays);
    for (uint32_t k = 0; k < n_rays; k++) {
        const float* r = rays12 + 12 * (size_t)k;
        list[k].origin = vec3(r[0], r[1], r[2]);
        list[k].direction = vec3(r[4], r[5], r[6]);
        list[k].probe_info = vec3(r[8], r[9], r[10]);
    }
    rays = list.data();
    probe_image_albedo.width = W; probe_image_albedo.height = H; probe_image_albedo.data = albedo; probe_image_albedo.f32 = albedo_f32;
    probe_image_distances.width = W; probe_image_distances.height = H; probe_image_distances.data = distances; probe_image_distances.f32 = nullptr;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t index = (size_t)W * y + x;
            if (index >= n_rays) continue;
            gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
            glsl_lookup_counter = 0;
            reinit_globals();
            shader_main();
            if (lookups) lookups[index] = glsl_lookup_counter;
        }
    rays = nullptr;
}
'''

HARNESS_PIXEL = r'''
namespace %(ns)s {
static void reinit_globals() { %(reinit)s }
}  // namespace %(ns)s

// One compute_pass.comp invocation per pixel of the floor(w/16) x floor(h/16) groups the
// reference dispatches (src/rvpt/rvpt.cpp:1139-1140).  cam20 = Camera::get_data().
extern "C" __attribute__((visibility("default")))
void %(entry)s(const RefSettings* s, const RefField* f, const float* cam20, const uint32_t* tex_albedo,
                      const uint32_t* tex_distances, int W, int H, uint32_t* frame, float* frame_f32, uint32_t* lookups)
{
    using namespace %(ns)s;
    render_settings.screen_width = s->screen_width; render_settings.screen_height = s->screen_height;
    render_settings.max_bounces = s->max_bounces; render_settings.camera_mode = s->camera_mode;
    render_settings.render_mode = s->render_mode; render_settings.scene = s->scene;
    render_settings.time = s->time; render_settings.visualize_probes = s->visualize_probes != 0;
    irradiance_field.probe_count = ivec3(f->probe_count[0], f->probe_count[1], f->probe_count[2]);
    irradiance_field.side_length = f->side_length; irradiance_field.hysteresis = f->hysteresis;
    irradiance_field.sqrt_rays_per_probe = f->sqrt_rays_per_probe;
    irradiance_field.field_origin = vec3(f->field_origin[0], f->field_origin[1], f->field_origin[2]);
    for (int c = 0; c < 4; c++) cam.matrix[c] = vec4(cam20[4 * c], cam20[4 * c + 1], cam20[4 * c + 2], cam20[4 * c + 3]);
    cam.params = vec4(cam20[16], cam20[17], cam20[18], cam20[19]);
    int w = s->screen_width, h = s->screen_height;
    result_image.width = w; result_image.height = h; result_image.data = frame; result_image.f32 = frame_f32;
    probe_image_albedo.width = W; probe_image_albedo.height = H; probe_image_albedo.data = const_cast<uint32_t*>(tex_albedo);
    probe_image_distances.width = W; probe_image_distances.height = H; probe_image_distances.data = const_cast<uint32_t*>(tex_distances);
    for (int y = 0; y < (h / 16) * 16; y++)
        for (int x = 0; x < (w / 16) * 16; x++) {
            gl_GlobalInvocationID = uvec3((uint)x, (uint)y, 0u);
            glsl_lookup_counter = 0;
            reinit_globals();
            shader_main();
            if (lookups) lookups[(size_t)y * w + x] = glsl_lookup_counter;
        }
}
'''


# Lines the reference has commented out, restored by removing the comment markers only.
RESTORE = {
    "hysteresis": (re.compile(
        r"/\*(\s*vec3 old_color = vec3\(imageLoad\(probe_image_albedo, texture_coords\)\);\s*"
        r"color = mix\(old_color, color, irradiance_field\.hysteresis\);)\*/"), r"\1"),
    "lights": (re.compile(r"//(update_lights\(\);)"), r"\1"),
    "chebyshev": (re.compile(r"//(weight \*= chebyshevWeight;)"), r"\1"),
}

LIGHT_TABLE = re.compile(r"^Light\s+(\w+)\s*\[(\w+\s*\[\s*\d+\s*\])\]\s*=\s*(\{.*?\})\s*;", re.S | re.M)


def build_unit(shader: str, ns: str, shader_dir: str, harness: str, entry: str = "", restore: str = "") -> tuple[str, list[str]]:
    seen: list[str] = []
    src = resolve_includes(os.path.join(shader_dir, shader), shader_dir, seen)
    if restore:
        pat, rep = RESTORE[restore]
        src, n = pat.subn(rep, src)
        if n != 1:
            raise SystemExit(f"{shader}: commented-out {restore} line not found exactly once ({n})")
    cpp = transpile(src)
    # lookup counter at the top of getBlockAt
    cpp, n = re.subn(r"(\bint\s+getBlockAt\s*\([^)]*\)\s*\{)", r"\1 glsl_count_lookup();", cpp, count=1)
    if n != 1:
        raise SystemExit(f"{shader}: getBlockAt not found")
    inits = file_scope_initialisers(cpp)
    reinit = " ".join(f"{name} = {expr};" for name, expr in inits)
    if restore == "lights":
        # update_lights() writes the file-scope light tables (structs.glsl:63-89); GLSL
        # re-initialises them for every invocation like any other global
        code_only = "".join(t if is_code else re.sub(r"[^\n]", " ", t) for is_code, t in split_comments(cpp))
        tables = LIGHT_TABLE.findall(code_only)
        if len(tables) != 3:
            raise SystemExit(f"{shader}: expected 3 light tables, found {len(tables)}")
        for name, size, init in tables:
            reinit += f" {{ Light init_[{size}] = {' '.join(init.split())}; for (int i_ = 0; i_ < {size}; i_++) {name}[i_] = init_[i_]; }}"
    body = f"namespace {ns} {{\n{cpp}\n}}  // namespace {ns}\n" + harness % {"reinit": reinit, "ns": ns, "entry": entry}
    return body, seen


HOST_HARNESS = r'''
// GENERATED by oracle/ref_glsl/build_ref.py — do not commit.
// The reference's HOST ray generator compiled from its own text: src/rvpt/probe.h is included where
// it lies, generate_samples / RVPT::generate_probe_rays are lines %(first)d-%(last)d of src/rvpt/rvpt.cpp and
// struct IrradianceField lines %(f0)d-%(f1)d of src/rvpt/rvpt.h, verbatim, against oracle/ref_glsl/glm_shim.
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <glm/glm.hpp>
#include "%(probe_h)s"

class RVPT
{
public:
    void generate_probe_rays();
%(field_struct)s
    IrradianceField ir;
    std::vector<ProbeRay> probe_rays;
    bool need_generate_probe_rays = true;
};

%(text)s

extern "C" __attribute__((visibility("default")))
uint32_t ref_generate_probe_rays(const int* probe_count, int side_length, int sqrt_rays_per_probe, const float* field_origin,
                                 int reseed, float* out12, uint32_t capacity)
{
    static_assert(sizeof(ProbeRay) == 48, "ProbeRay is 48 bytes");
    if (reseed) srand(1);   // the state of a process that never called srand
    RVPT r;
    r.ir.probe_count = glm::ivec3(probe_count[0], probe_count[1], probe_count[2]);
    r.ir.side_length = side_length;
    r.ir.sqrt_rays_per_probe = sqrt_rays_per_probe;
    r.ir.field_origin = glm::vec3(field_origin[0], field_origin[1], field_origin[2]);
    r.generate_probe_rays();
    uint32_t n = (uint32_t)r.probe_rays.size();
    if (out12 && n <= capacity) {
        memset(out12, 0, (size_t)n * 48);
        for (uint32_t k = 0; k < n; k++) {
            const ProbeRay& p = r.probe_rays[k];
            float* o = out12 + 12 * (size_t)k;
            o[0] = p.origin.x; o[1] = p.origin.y; o[2] = p.origin.z;
            o[4] = p.direction.x; o[5] = p.direction.y; o[6] = p.direction.z;
            o[8] = p.probe_info.x; o[9] = p.probe_info.y; o[10] = p.probe_info.z;
        }
    }
    return n;
}
'''


def build_host_unit(reference: str) -> str:
    """The reference's host ray generator (rvpt.cpp: #define PI ... end of RVPT::generate_probe_rays)."""
    cpp = open(os.path.join(reference, "src", "rvpt", "rvpt.cpp"), encoding="utf-8", errors="replace").read().split("\n")
    first = next(i for i, l in enumerate(cpp) if l.startswith("#define PI"))
    start_fn = next(i for i, l in enumerate(cpp) if l.startswith("void RVPT::generate_probe_rays()"))
    depth, last = 0, None
    for i in range(start_fn, len(cpp)):
        depth += cpp[i].count("{") - cpp[i].count("}")
        if depth == 0 and "}" in cpp[i] and i > start_fn:
            last = i
            break
    if last is None:
        raise SystemExit("rvpt.cpp: end of RVPT::generate_probe_rays not found")
    hdr = open(os.path.join(reference, "src", "rvpt", "rvpt.h"), encoding="utf-8", errors="replace").read().split("\n")
    f0 = next(i for i, l in enumerate(hdr) if l.strip() == "struct IrradianceField")
    f1 = next(i for i in range(f0, len(hdr)) if hdr[i].strip() == "};")
    return HOST_HARNESS % {"first": first + 1, "last": last + 1, "f0": f0 + 1, "f1": f1 + 1,
                           "probe_h": os.path.join(reference, "src", "rvpt", "probe.h"),
                           "field_struct": "\n".join(hdr[f0:f1 + 1]), "text": "\n".join(cpp[first:last + 1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    shader_dir = os.path.join(args.reference, "assets", "shaders")
    if not os.path.isdir(shader_dir):
        print(f"build_ref: {shader_dir} not present — keeping any prebuilt oracle/_ref as is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    head = '#include <vector>\n#include "../ref_glsl/glsl_shim.h"\n' + HARNESS_COMMON
    units = []
    for shader, ns, harness, entry, restore in (
            ("probe_pass.comp", "probe_pass", HARNESS_PROBE, "ref_probe_pass", ""),
            ("probe_pass.comp", "probe_pass_hysteresis", HARNESS_PROBE, "ref_probe_pass_hysteresis", "hysteresis"),
            ("probe_pass.comp", "probe_pass_lights", HARNESS_PROBE, "ref_probe_pass_lights", "lights"),
            ("compute_pass.comp", "compute_pass", HARNESS_PIXEL, "ref_compute_pass", ""),
            ("compute_pass.comp", "compute_pass_lights", HARNESS_PIXEL, "ref_compute_pass_lights", "lights"),
            ("compute_pass.comp", "compute_pass_chebyshev", HARNESS_PIXEL, "ref_compute_pass_chebyshev", "chebyshev")):
        body, seen = build_unit(shader, ns, shader_dir, harness, entry, restore)
        path = os.path.join(OUT, ns + "_ref.cpp")
        with open(path, "w") as f:
            f.write(f"// GENERATED by oracle/ref_glsl/build_ref.py from {shader} + {', '.join(seen)} — do not commit\n")
            f.write(head + body)
        units.append(path)
    with open(os.path.join(OUT, "host_rays_ref.cpp"), "w") as f:
        f.write(build_host_unit(args.reference))
    with open(os.path.join(OUT, "globals_ref.cpp"), "w") as f:
        f.write('#include "../ref_glsl/glsl_shim.h"\nuvec3 gl_GlobalInvocationID;\nuint32_t glsl_lookup_counter = 0;\n')
    units.append(os.path.join(OUT, "globals_ref.cpp"))
    subprocess.check_call(["make", "-s", "-C", ORACLE])
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # the host ray generator is its own object: it must see the glm stand-in, not the GLSL shim
    host_obj = os.path.join(OUT, "host_rays_ref.o")
    host_cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-c", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-w",
                "-I" + os.path.join(HERE, "glm_shim"), "-o", host_obj, os.path.join(OUT, "host_rays_ref.cpp")]
    print(" ".join(host_cmd))
    subprocess.check_call(host_cmd)
    units.append(host_obj)
    cmd = [cxx, "-O2", "-std=c++20", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
           "-fpermissive", "-w", "-o", os.path.join(OUT, "libddgi_ref.so"), *units,
           "-L" + ORACLE, "-lddgi_oracle", "-Wl,-rpath,$ORIGIN/..", "-lm"]
    print(" ".join(cmd))
    subprocess.check_call(cmd)
    print("built", os.path.join(OUT, "libddgi_ref.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
