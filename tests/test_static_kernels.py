"""What the built library says about its kernels without a GPU (not gpu): the cubins are sm_100a, the Q1 kernel moves its
columns with TMA bulk copies, and the two hot kernels stay inside the register budgets their launch shapes assume
(k_scan_agg_small: one 480-thread CTA per SM; k_probe_chain: four 256-thread CTAs per SM).  A build-time guard for the
numbers in DESIGN.md 3 / profiles/r01h_static_kernels.txt - a kernel that silently grows past its budget loses occupancy
long before a parity test notices."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "cloudberry_b200", "libcbgpu.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(SO) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")


@pytest.fixture(scope="module")
def res_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", SO], capture_output=True, text=True, timeout=300).stdout
    table, fn = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and fn:
            table[fn] = tuple(int(x) for x in m.groups())
    return out, table


def test_cubins_are_sm_100a(res_usage):
    out, table = res_usage
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs
    assert len(table) > 40


def test_hot_kernels_stay_inside_their_register_budgets(res_usage):
    _, table = res_usage
    scan = {k: v for k, v in table.items() if "k_scan_agg_small" in k}
    assert scan
    four_groups = [v for k, v in scan.items() if "ILi4E" in k]
    assert four_groups and all(reg <= 128 for reg, _, _ in four_groups)         # 480 threads x 128 = 61 440 of 65 536 registers
    (reg, stack, shared), = [v for k, v in table.items() if "k_probe_chain" in k]
    assert reg <= 64                                                               # 4 CTAs x 256 threads x 64 = 65 536
    assert shared <= 56 * 1024                                                     # four CTAs' static shared memory per SM


def test_q1_kernel_uses_tma_bulk_copies():
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", "_Z16k_scan_agg_smallILi4ELi55ELb1ELi1EEv14SmallAggParams", SO],
                          capture_output=True, text=True, timeout=300).stdout
    assert "UBLKCP" in sass, "no cp.async.bulk in the Q1 kernel"
    assert "SYNCS" in sass, "no mbarrier traffic in the Q1 kernel"
    assert len(re.findall(r"\bLDG\.E\.(64|128)", sass)) == 0                       # columns come through shared memory, not LDG
