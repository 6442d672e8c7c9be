"""Library-level multi-GPU (wg_multi: instance i -> device i mod G, one NCCL all-reduce of the statistics)."""
import ctypes as C

import numpy as np
import pytest


def test_multi_create_without_a_device_fails_loudly():
    """No CPU fallback: on a box without CUDA devices wg_multi_create reports WG_ERR_NO_DEVICE (-1)."""
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    if lib.wg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.wg_multi_create(0, C.byref(h)) == -1 and not h.value
    assert lib.wg_multi_size(None) == 0 and lib.wg_multi_nccl_version(None) == 0


@pytest.mark.gpu
def test_gpu_sharded_sweep_equals_the_single_context_run(ctx):
    """wg_multi_herdt_mpc_sweep over every visible device (one on the test box, 2/4/8 under `gpurun --gpus N`): the reduced
    statistics equal those of the same instances run through one context, every device got its round-robin share, and the
    reduction went over NCCL."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(4)
    B, periods = 6000, 25
    v = np.column_stack([rng.uniform(-0.2, 0.3, B), rng.uniform(-0.15, 0.15, B), rng.uniform(-0.2, 0.2, B)])
    ctx.herdt_set_params(); ctx.herdt_mpc_set_params()
    st = ctx.herdt_mpc_init(B)
    ctx.herdt_mpc_run(st, periods, vel_ref=v)
    m = wg.MultiContext(0)
    try:
        m.herdt_set_params()
        r = m.herdt_mpc_sweep(v, periods, chunk=10)
    finally:
        m.close()
    G = r["devices"]
    assert G == wg.device_count() and r["device_instances"] == [len(range(k, B, G)) for k in range(G)]
    assert r["qp_solves"] == st["qp_count"].sum() == B * periods
    assert r["failures"] == st["fail_count"].sum() == 0
    assert r["iterations"] == st["iterations_total"].sum()
    assert r["still_online"] == B and r["seconds"] > 0 and r["seconds"] == pytest.approx(max(r["device_ms"]) * 1e-3)
    assert r["reduced_by_nccl"] and r["nccl_version"] > 20000, r
    print(f"wg_multi: {G} device(s), NCCL {r['nccl_version']}, {r['qp_solves'] / r['seconds'] / 1e6:.1f} M closed-loop solves/s")
