"""Static checks on the compiled library (no GPU): the tcgen05 attention kernels carry tensor-core / TMA instructions and
issue them from an ELECTED lane.  From a thread-index branch (`if (tid == 0)`) nvcc wraps every tcgen05.mma / TMA instruction
in an ELECT / R2UR / BRA.U.ANY loop over the active lanes (~55-95 cycles per MMA, profiles/r2_ncu_attn_bwd_pair.md,
tools/micro/umma_rate.cu); behind `elect.sync` they are issued back to back.  This guards that finding."""
import re
import shutil
import subprocess

import pytest

from mtvaf_b200 import lib as L

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
KERNELS = ("attn_fwd_tc_kernel", "attn_fwd_tc_pair_kernel", "attn_bwd_pipe_kernel", "attn_bwd_pair_kernel",
           "attn_bwd_long_kernel", "attn_bwd_tc_kernel", "pairwise_gram_tc_kernel")


@pytest.fixture(scope="module")
def sass():
    try:
        out = subprocess.run([CUOBJDUMP, "-sass", L.LIB_PATH], capture_output=True, text=True, timeout=600)
    except (FileNotFoundError, subprocess.TimeoutExpired) as e:      # pragma: no cover
        pytest.skip("cuobjdump unavailable: %r" % (e,))
    if out.returncode != 0:                                          # pragma: no cover
        pytest.skip("cuobjdump failed: " + out.stderr[-200:])
    per, cur = {}, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = []
        elif cur is not None:
            per[cur].append(line)
    return per


@pytest.mark.parametrize("kernel", KERNELS)
def test_tcgen05_attention_kernels_issue_from_an_elected_lane(sass, kernel):
    fns = [k for k in sass if kernel in k]
    assert fns, "kernel %s not found in %s" % (kernel, L.LIB_PATH)
    for fn in fns:
        text = "\n".join(sass[fn])
        assert "UTCHMMA" in text, fn + ": no tcgen05.mma"
        if kernel != "pairwise_gram_tc_kernel":          # (the Gram kernel splits fp32 rows into bf16 hi / lo itself: no TMA)
            assert "UTMALDG" in text, fn + ": no TMA load"
        assert "LDTM" in text, fn + ": no tcgen05.ld"
        assert "BRA.U.ANY" not in text, fn + ": tcgen05 / TMA issue serialised over lanes (thread-index branch?)"
