"""Several engines in ONE process behind the C ABI (relxill_b200_init_devices, include/relxill_b200.h): a host-buffer
batch is sharded over all visible devices, every device copies its rows straight into the caller's array.  Runs last
(it re-initialises the runtime); on a one-GPU box it still goes through the multi-engine entry points."""
import numpy as np
import pytest

from common import default_grid, sample_params, walker_ball


@pytest.mark.gpu
def test_batch_sharded_over_all_devices_in_one_process(rx, table_dir):
    e = default_grid(800)
    cases = [("relxilllp", walker_ball("relxilllp", 257, seed=11)), ("relxillCp", sample_params("relxillCp", 130, seed=12)),
             ("relline", sample_params("relline", 64, seed=13))]
    rx.set_cache(False)
    want = [rx.batch_eval(m, e, P, return_status=True) for m, P in cases]
    rx.shutdown()
    try:
        n = rx.init_devices(table_dir, 0)          # all visible devices
        assert n >= 1 and rx.num_devices() == n
        for interleave in (False, True):
            rx.set_sharding(interleave)
            for (m, P), (f0, s0) in zip(cases, want):
                f, s = rx.batch_eval(m, e, P, return_status=True)
                np.testing.assert_array_equal(s, s0)
                np.testing.assert_array_equal(f, f0)   # a spectrum does not depend on the device or the shard it ran in
        # the XSPEC symbols go through the same engines
        m, P = cases[0]
        np.testing.assert_array_equal(rx.lmod(m, e, P[3]), want[0][0][3])
    finally:
        rx.set_sharding(False)
        rx.shutdown()
        rx.init(table_dir)
        rx.set_cache(True)
