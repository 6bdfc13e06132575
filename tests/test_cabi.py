"""The C-ABI library loads on a CPU-only box and exports every symbol include/svoxb.h declares. No compute calls."""
import ctypes
import os
import re

import svox_t_b200.csrc as C

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "svoxb.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SVOXB_API\s+[\w\s\*]+?\b(svoxb_\w+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ("svoxb_query", "svoxb_render_rays_fwd", "svoxb_render_rays_bwd", "svoxb_render_image_fwd",
                 "svoxb_render_image_bwd", "svoxb_render_depth", "svoxb_construct_tree", "svoxb_warp_vertices",
                 "svoxb_p2v", "svoxb_build_octree_count", "svoxb_build_octree_emit", "svoxb_accel_create",
                 "svoxb_query_f64", "svoxb_render_rays_fwd_f64", "svoxb_render_rays_bwd_f64",
                 "svoxb_render_image_fwd_f64", "svoxb_render_image_bwd_f64", "svoxb_render_depth_f64"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/svoxb.h but not exported by libsvoxb.so"
    assert sorted(C.SYMBOLS) == declared_symbols(), "python prototypes out of sync with include/svoxb.h"
    assert lib.svoxb_abi_version() == 10


def test_struct_layouts_match_header():
    # svoxb_render_options: 11 four-byte fields in the reference's order (data_spec.hpp:129-145)
    assert ctypes.sizeof(C._COptions) == 44
    assert [f[0] for f in C._COptions._fields_] == ["step_size", "background_brightness", "format", "basis_dim",
                                                    "ndc_width", "ndc_height", "ndc_focal", "min_comp", "max_comp",
                                                    "sigma_thresh", "stop_thresh"]
    assert ctypes.sizeof(C._CTree) == 144 and ctypes.sizeof(C._CCamera) == 32


def test_ctypes_structs_match_the_header_as_gcc_lays_it_out(tmp_path):
    """Compile include/svoxb.h with gcc and compare sizeof / offsetof of every struct field with the ctypes mirrors."""
    import subprocess
    structs = {"svoxb_tree": C._CTree, "svoxb_render_options": C._COptions, "svoxb_camera": C._CCamera,
               "svoxb_tree_f64": C._CTree64, "svoxb_camera_f64": C._CCamera64}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.abspath(HEADER)}"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_bad_arguments_return_error_codes_not_crashes():
    lib = C.load_library()
    rc = lib.svoxb_render_rays_fwd(None, None, None, None, 0, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.svoxb_last_error()
    assert lib.svoxb_accel_describe(None, None, None, None) == -1
    assert lib.svoxb_render_rays_fwd_f64(None, None, None, 0, None, None, None, None) == -1
    assert b"NULL" in lib.svoxb_last_error()
    assert lib.svoxb_launch_count() == 0 or lib.svoxb_launch_count() > 0
