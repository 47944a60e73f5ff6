"""The two FP64 micro-benchmarks behind bench.py's roofline.fp64 block: plain DFMA chains and the tensor-core path
(mma.sync m8n8k4 f64).  They only have to run, report the flops they executed and write finite results."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_fp64_probes_run_and_report_their_flops():
    import torch
    import reyna_b200
    L = reyna_b200.lib()
    scratch = torch.zeros(148 * 8 * 256, dtype=torch.float64, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fl = C.c_double(0.0)
    assert L.reyna_b200_fma_probe(100, scratch.data_ptr(), C.byref(fl), sp) == 0
    assert fl.value == 2.0 * 8 * 100 * 148 * 8 * 256
    torch.cuda.synchronize()
    assert bool(torch.isfinite(scratch).all())
    assert L.reyna_b200_dmma_probe(100, scratch.data_ptr(), C.byref(fl), sp) == 0
    assert fl.value == 2.0 * 8 * 8 * 4 * 8 * 100 * 148 * 8 * 8            # 8 tiles of m8n8k4 per warp, 8 warps per CTA
    torch.cuda.synchronize()
    assert bool(torch.isfinite(scratch).all()) and float(scratch.abs().max()) > 0.0
    assert L.reyna_b200_dmma_probe(0, scratch.data_ptr(), C.byref(fl), sp) != 0      # bad argument is an error, not a no-op
