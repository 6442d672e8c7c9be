"""CPU checks of the drop-in boundary: the C ABI header, the library's exports, struct layouts, host-only entry points,
and the no-CPU-fallback rule.  No compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "walkgen_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_symbol_the_header_declares():
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/walkgen_b200.h but not exported"
    assert sorted(_capi.SIGNATURES) == names, set(names) ^ set(_capi.SIGNATURES)


def test_header_is_plain_c_and_struct_layouts_match_the_python_mirrors(tmp_path):
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import _capi
    prog = tmp_path / "sizes.c"
    structs = ["wg_preview_gains_t", "wg_herdt_params", "wg_herdt_qp_input", "wg_herdt_qp_output", "wg_herdt_mpc_params",
               "wg_herdt_foot_sample", "wg_herdt_tick", "wg_herdt_mpc_state", "wg_herdt_mpc_step", "wg_pldp_state",
               "wg_pldp_info", "wg_pldp_batch", "wg_rel_step", "wg_foot_sample", "wg_zmpdisc_params", "wg_lci",
               "wg_dimitrov_params", "wg_dimitrov_period"]
    body = "\n".join(f'  printf("{s} %zu\\n", sizeof({s}));' for s in structs)
    prog.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(prog)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    sizes = {k: int(v) for k, v in out.items()}
    assert sizes["wg_herdt_qp_input"] == wg.QP_INPUT_DTYPE.itemsize == 784
    assert sizes["wg_herdt_qp_output"] == wg.QP_OUTPUT_DTYPE.itemsize == 960
    assert sizes["wg_herdt_foot_sample"] == wg.FOOT_DTYPE.itemsize
    assert sizes["wg_herdt_tick"] == wg.TICK_DTYPE.itemsize == 256
    assert sizes["wg_herdt_mpc_state"] == wg.MPC_STATE_DTYPE.itemsize
    assert sizes["wg_herdt_mpc_step"] == wg.MPC_STEP_DTYPE.itemsize
    assert sizes["wg_pldp_state"] == wg.PLDP_STATE_DTYPE.itemsize
    assert sizes["wg_pldp_info"] == wg.PLDP_INFO_DTYPE.itemsize
    assert sizes["wg_preview_gains_t"] == C.sizeof(_capi.PreviewGains)
    assert sizes["wg_herdt_params"] == C.sizeof(_capi.HerdtParams)
    assert sizes["wg_herdt_mpc_params"] == C.sizeof(_capi.HerdtMpcParams)
    assert sizes["wg_pldp_batch"] == C.sizeof(_capi.PldpBatch)
    assert sizes["wg_rel_step"] == wg.REL_STEP_DTYPE.itemsize == 48
    assert sizes["wg_foot_sample"] == wg.KAJITA_FOOT_DTYPE.itemsize == 48
    assert sizes["wg_zmpdisc_params"] == C.sizeof(_capi.ZmpDiscParams)
    assert sizes["wg_lci"] == wg.LCI_DTYPE.itemsize == 272
    assert sizes["wg_dimitrov_period"] == wg.DIMITROV_PERIOD_DTYPE.itemsize == 224
    assert sizes["wg_dimitrov_params"] == C.sizeof(_capi.DimitrovParams)
    import dimitrov_oracle as do
    assert C.sizeof(do.Params) == sizes["wg_dimitrov_params"] and do.LCI.itemsize == 272 and do.PERIOD.itemsize == 224


def test_host_only_entry_points_work_without_a_gpu():
    import jrl_walkgen_b200 as wg
    p = wg.herdt_default_params(0.25, 0.14)
    # ZMPVelocityReferencedQP.cpp:63,103,116-118; FootHalfSize.cpp:62-68 with the 0.04 margins
    assert (p.T, p.com_height, p.w_jerk, p.w_vel, p.w_cop) == (0.1, 0.814, 1e-5, 1.0, 1e-6)
    assert abs(p.cop_half_x - 0.085) < 1e-15 and abs(p.cop_half_y - 0.03) < 1e-15
    assert list(p.foot_hull_x) == [-0.28, -0.2, 0.0, 0.2, 0.28]
    m = wg.herdt_mpc_default_params()
    assert (m.Ts, m.time_buffer, m.step_period, m.dsss_period, m.nb_steps_ssds) == (0.005, 0.04, 0.8, 0.8, 2)
    g = wg.preview_gains(0.005, 1.6, 0.814, wg.MODE_WITHOUT_INITIALPOS)
    assert g.NL == 320 and abs(g.Ks - 618.7) / 618.7 < 2e-4     # src/data/PreviewControlParameters.ini
    from jrl_walkgen_b200 import _capi
    assert _capi.load().wg_version() > 0
    # Dimitrov: defaults of ZMPConstrainedQPFastFormulation's ctor (:79-96) and the loop bound of :1190-1194 on the accumulated
    # 5 ms clock, against the oracle's restatement (host arithmetic only)
    import dimitrov_oracle as do
    d = wg.dimitrov_default_params()
    assert (d.T, d.sampling_period, d.com_height, d.alpha, d.beta, d.constraint_x, d.constraint_y) == (0.1, 0.005, 0.80, 200.0, 1000.0, 0.04, 0.04)
    assert (d.cold_restart, d.merge_duplicate_rows) == (0, 0)
    for n in (1, 2, 321, 322, 1602, 4002, 4003, 9999, 20001):
        assert _capi.load().wg_dimitrov_period_count(C.byref(d), n) == do.period_count(n), n
    z = wg.zmpdisc_default_params()
    steps = np.zeros(3, dtype=wg.REL_STEP_DTYPE); steps["ss_time"], steps["ds_time"] = 0.78, 0.02
    assert _capi.load().wg_zmpdisc_sample_count(C.byref(z), 3, steps.ctypes.data) == 640 + 2 * 160 + 2 + 960


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run: it never falls back to a CPU implementation."""
    import jrl_walkgen_b200 as wg
    if wg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    from jrl_walkgen_b200 import _capi
    assert _capi.load().wg_ctx_create(0, C.byref(h)) == _capi.WG_ERR_NO_DEVICE and not h.value
    with pytest.raises(wg.WalkgenError) as ei:
        wg.Context(0)
    assert ei.value.code == _capi.WG_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "jrl_walkgen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "herdt_oracle" not in txt, os.path.join(dirpath, f)
    lib = os.path.join(pkg, "libwalkgen_b200.so")
    syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


@pytest.mark.gpu
def test_gpu_memset_device_fill_kernel_and_fallback():
    """wg_memset_device: 16-byte aligned fills run the library's own grid-stride kernel, anything else cudaMemsetAsync; both must
    set exactly the bytes asked for (byte value replicated) and nothing around them."""
    import numpy as np
    import jrl_walkgen_b200 as wg
    ctx = wg.Context(0)
    try:
        n = 1 << 20
        buf = ctx.to_device(np.full(n, 0x11, dtype=np.uint8))
        for off, cnt, val in ((0, n, 0), (4096, 65536, 0xAB), (16, 48, 0x7F), (8, 100, 0x01), (3, 5, 0xFF), (32, 0, 0x55)):
            ref = np.full(n, 0x11, dtype=np.uint8)
            buf.upload(ref)
            ref[off:off + cnt] = val
            ctx._check(ctx.lib.wg_memset_device(ctx.h, buf.ptr + off, val, cnt))
            got = buf.download(np.uint8, (n,))
            assert np.array_equal(got, ref), (off, cnt, val)
        buf.free()
    finally:
        ctx.close()
