"""Static checks on the compiled library (no GPU needed): the hot kernels really are tcgen05 / TMA kernels for
sm_100a, and the issue loops of the two kernels that were found issue-bound stay free of the ELECT / R2UR waterfall
loops the compiler emits around tcgen05 / TMA instructions inside `if (lane == 0)` regions
(profiles/r01_kernel_pool_investigation.md)."""
import re
import shutil
import subprocess

import pytest

from matchmaker_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    try:
        out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    except (FileNotFoundError, subprocess.TimeoutExpired) as e:
        pytest.skip(f"cuobjdump unavailable: {e}")
    if out.returncode != 0:
        pytest.skip("cuobjdump failed: " + out.stderr[-200:])
    funcs, name = {}, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    assert "sm_100a" in out.stdout or "SM100" in out.stdout.upper() or funcs, "no sm_100a code in the library"
    return {k: "\n".join(v) for k, v in funcs.items()}


def _kernels(sass, needle):
    ks = {k: v for k, v in sass.items() if needle in k}
    assert ks, f"no kernel matching {needle!r} in the library"
    return ks


@pytest.mark.parametrize("needle", ["maxsim_qm_kernel", "kernel_pool_ts_kernel", "flat_ip_tc_kernel",
                                    "maxsim_tc_kernel", "kernel_pool_bwd_tc_kernel", "tkl_ts_kernel"])
def test_tensor_core_kernels_use_tcgen05_and_tma(sass, needle):
    for name, text in _kernels(sass, needle).items():
        assert "UTCHMMA" in text, f"{name}: no tcgen05.mma (UTCHMMA) in the SASS"
        assert "UTMALDG" in text, f"{name}: no TMA tensor load (UTMALDG) in the SASS"
        assert "UTCBAR" in text, f"{name}: no tcgen05.commit (UTCBAR) in the SASS"
        assert re.search(r"\bLDTM\b|LDTM\.", text), f"{name}: accumulators are never read back from tensor memory (LDTM)"


def test_kernel_pool_ts_feeds_the_mma_from_tensor_memory(sass):
    for name, text in _kernels(sass, "kernel_pool_ts_kernel").items():
        assert re.search(r"\bSTTM\b|STTM\.", text), f"{name}: no tcgen05.st (STTM): the document operand is not written to TMEM"
        assert "USETMAXREG" in text, f"{name}: setmaxnreg is missing"


@pytest.mark.parametrize("needle", ["kernel_pool_ts_kernel", "flat_ip_tc_kernel"])
def test_mma_issue_is_uniform(sass, needle):
    """The waterfall pattern is `UTCHMMA ... ; @P0 BRA.U.ANY <back>`: no MMA of these kernels may be followed by one."""
    for name, text in _kernels(sass, needle).items():
        lines = [l for l in text.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        for i, l in enumerate(lines):
            if "UTCHMMA" in l:
                nxt = " ".join(lines[i + 1:i + 3])
                assert "BRA.U.ANY" not in nxt, f"{name}: tcgen05.mma inside a waterfall loop (issue it under elect.sync)"


def test_flat_ip_uses_multicast_in_the_cluster_instantiations(sass):
    ks = _kernels(sass, "flat_ip_tc_kernel")
    multi = [k for k, v in ks.items() if "UTMALDG" in v and ".MULTICAST" in v.upper()]
    assert multi, "no flat-IP instantiation issues a multicast TMA load"


def test_kernel_pool_backward_is_a_tensor_core_kernel_with_tma_stores(sass):
    """Both contractions of the backward as UMMAs (A once from tensor memory, once from shared memory), G written to
    tensor memory, gradients leaving through TMA stores."""
    for name, text in _kernels(sass, "kernel_pool_bwd_tc_kernel").items():
        assert len(re.findall(r"UTCHMMA", text)) >= 2, f"{name}: expected the two UMMA chains"
        assert re.search(r"\bSTTM\b|STTM\.", text), f"{name}: no tcgen05.st (STTM): G is not written to TMEM"
        assert "UTMASTG" in text, f"{name}: no TMA tensor store (UTMASTG)"
        assert "USETMAXREG" in text, f"{name}: setmaxnreg is missing"

