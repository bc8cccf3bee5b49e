"""CPU: resource footprint of the kernels that run NEXT TO a blend CTA in the fused path.
An SM that runs blend_ws_kernel (128 registers x 384 threads, 231 KB of shared memory) has 16 K registers and 1.4 KB of shared
memory left: room for one small CTA.  A kernel of the auxiliary-stream chain that does not fit waits for the blend to end and
stalls the whole chain behind it (round 2 lost 3 % of the step to an 80-register resize kernel), so the limits are pinned here."""
import re
import shutil
import subprocess

import pytest

from simplepanorama_b200 import _lib as L

CHAIN = ["resize_tables_kernel", "resize_activity_kernel", "resize_linear_u8_kernel", "warp_tables_kernelILi0E", "warp_tables_kernelILi1E", "warp_kernelILi0E",
         "warp_kernelILi1E", "ccl_init_kernel", "ccl_merge_kernel", "resolve_bits_kernel", "erode_bits_kernel", "plan_init_kernel",
         "activity_kernelILi32E", "activity_kernelILi16E", "plan_kernelILi32E", "plan_kernelILi16E", "normalise_kernel"]
THREADS = 256                      # largest CTA any of them is launched with
FREE_REGS = 65536 - 128 * 384      # next to blend_ws_kernel<B, SW, 256>
FREE_SMEM = 233472 - 232064        # SM shared memory - (blend dynamic 231040 + 1 KB system reserve)


def _usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "--dump-resource-usage", L.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except Exception:
        pytest.skip("cuobjdump not available")
    use, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) .*SHARED:(\d+)", line)
        if m and name:
            use[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    if not use:
        pytest.skip("no resource usage in cuobjdump output")
    return use


def test_aux_chain_kernels_fit_next_to_a_blend_cta(spano_lib):
    use = _usage()
    for key in CHAIN:
        hits = [(n, v) for n, v in use.items() if key in n and "tma" not in n]
        assert hits, f"kernel {key} not found in libspano.so"
        for n, (regs, smem) in hits:
            alloc = (regs + 7) // 8 * 8 * THREADS           # registers are allocated in units of 8 per thread
            assert alloc <= FREE_REGS, f"{key}: {regs} registers x {THREADS} threads does not fit next to a blend CTA"
            assert smem <= FREE_SMEM, f"{key}: {smem} B of shared memory (incl. the 1 KB reserve) does not fit next to a blend CTA"


def test_blend_kernel_leaves_room(spano_lib):
    use = _usage()
    ws = [(n, v) for n, v in use.items() if "blend_ws_kernelILi6ELi32ELi256ELb0" in n]
    assert ws and ws[0][1][0] == 128, "blend_ws_kernel<6,32,256> must be capped at 128 registers (the launch allocates 128 x 384)"
