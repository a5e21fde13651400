"""The C-ABI library loads without a GPU and exports every symbol include/ddgi.h declares;
the PODs have the reference's byte layouts (src/rvpt/rvpt.h:70-90, src/rvpt/probe.h:5-20).
No compute entry point is called here."""
import ctypes as C
import os
import re

import ddgi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = ddgi_b200.capi


def _declared():
    text = open(os.path.join(ROOT, "include", "ddgi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"DDGI_API\s+[\w\s\*]+?\b(ddgi_\w+)\s*\(", text)


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 35 and len(set(names)) == len(names)
    lib = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ddgi.h but not exported"
        assert n in capi.PROTOTYPES, f"{n} has no ctypes prototype in capi.py"
    assert set(capi.PROTOTYPES) == set(names)


def test_pod_layouts_match_the_reference():
    assert C.sizeof(capi.RenderSettings) == 32
    assert [f[0] for f in capi.RenderSettings._fields_] == [
        "screen_width", "screen_height", "max_bounces", "camera_mode", "render_mode", "scene", "time", "visualize_probes"]
    f = capi.IrradianceField
    assert C.sizeof(f) == 48
    assert (f.probe_count.offset, f.side_length.offset, f.hysteresis.offset, f.sqrt_rays_per_probe.offset,
            f.field_origin.offset, f.visualize.offset) == (0, 12, 16, 20, 32, 44)
    r = capi.ProbeRay
    assert C.sizeof(r) == 48 and (r.origin.offset, r.direction.offset, r.probe_info.offset) == (0, 16, 32)
    assert C.sizeof(capi.Light) == 28


def test_version_and_no_cpu_fallback():
    lib = capi.load()
    assert b"sm_100a" in lib.ddgi_version()
    import torch

    if not torch.cuda.is_available():
        ctx = C.c_void_p()
        assert lib.ddgi_create(C.byref(ctx), 0) == capi.E_CUDA  # fails loudly, no CPU path
        assert not ctx.value


def test_light_tables_match_structs_glsl():
    """assets/shaders/structs.glsl:61-89: num_lights = {1,1,2}."""
    lib = capi.load()
    arr = (capi.Light * capi.MAX_LIGHTS)()
    n = C.c_int32()
    want = {0: (1, 100.0, (4.0, 17.5, 8.5)), 1: (1, 15.0, (0.0, 8.0, 13.0))}
    for scene, (cnt, inten, pos) in want.items():
        assert lib.ddgi_default_lights(scene, arr, C.byref(n)) == 0
        assert n.value == cnt and arr[0].intensity == inten and tuple(arr[0].pos) == pos
    assert lib.ddgi_default_lights(2, arr, C.byref(n)) == 0 and n.value == 2
    assert lib.ddgi_default_lights(7, arr, C.byref(n)) == capi.E_INVALID


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dynamic-diffuse-global-illumination-minecraft_b200")
    bad = re.compile(r"^\s*(import|from)\s+oracle\b|ddgi_oracle|libddgi_ref|oracle/_ref|#include\s+\"[^\"]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert not bad.search(text), f"{fn} reaches into oracle/"


def test_header_is_plain_c_and_the_cpp_mirror_compiles(tmp_path):
    """include/ddgi.h must be usable from C (C99, -pedantic) — it is the FFI boundary — and
    include/rvpt_ddgi.hpp from C++17 without warnings; both link against the in-tree library."""
    import subprocess

    inc = os.path.join(ROOT, "include")
    pkg = os.path.dirname(capi.LIB_PATH)
    c_src = tmp_path / "use_ddgi.c"
    c_src.write_text(
        '#include "ddgi.h"\n'
        "#include <stdio.h>\n"
        "int main(void) {\n"
        "    ddgi_light l[DDGI_MAX_LIGHTS], moved[DDGI_MAX_LIGHTS];\n"
        "    int32_t n = 0;\n"
        "    ddgi_ctx* ctx = 0;\n"
        "    if (sizeof(ddgi_render_settings) != 32 || sizeof(ddgi_irradiance_field) != 48 || sizeof(ddgi_probe_ray) != 48) return 2;\n"
        "    if (ddgi_default_lights(0, l, &n) != DDGI_OK || n != 1) return 3;\n"
        "    if (ddgi_update_lights(0, 74.0f, l, n, moved) != DDGI_OK) return 4;\n"
        '    printf("%s %d %.6f\\n", ddgi_version(), ddgi_create(&ctx, 0), (double)moved[0].pos[2]);\n'
        "    if (ctx) ddgi_destroy(ctx);\n"
        "    return 0;\n"
        "}\n")
    exe = tmp_path / "use_ddgi"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + inc, str(c_src), "-o", str(exe),
                           "-L" + pkg, "-lddgi_b200", "-Wl,-rpath," + pkg])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out[1] == "sm_100a" and float(out[3]) != 8.5          # the cave light moved with time
    import torch

    assert int(out[2]) == (capi.OK if torch.cuda.is_available() else capi.E_CUDA)
    cpp_src = tmp_path / "use_mirror.cpp"
    cpp_src.write_text('#include "rvpt_ddgi.hpp"\nint main() { ddgi::RVPT r(64, 64); return r.render_settings.max_bounces == 8 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + inc, str(cpp_src), "-o", str(tmp_path / "use_mirror"),
                           "-L" + pkg, "-lddgi_b200", "-Wl,-rpath," + pkg])
    assert subprocess.call([str(tmp_path / "use_mirror")]) == 0
