"""tcgen05.mma dispatch-rate microbenchmark (one idle SM): cycles per MMA by width, operand source and TMEM-read contention."""
import ctypes
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402,F401
import skin_sm3_b200 as sm3  # noqa: E402

lib = sm3.lib()
print("   n  A-src  ldtm_warps  count | issue cyc/MMA  total cyc/MMA   (floor by the guide: 128*n/256 = n/2)  ld round trips")
for n in (64, 128, 256):
    for a_tmem in (0, 1):
        for lw in (0, 4, 8):
            for count in (64, 512):
                buf = (ctypes.c_longlong * 16)()
                rc = lib.sm3_debug_umma_rate(n, a_tmem, count, lw, buf)
                assert rc == 0, sm3._lib.last_error()
                print(f"{n:4d}  {'TMEM' if a_tmem else 'smem'}  {lw:10d}  {count:5d} | {buf[0] / count:13.1f}  {buf[1] / count:13.1f}"
                      f"   {n / 2:5.0f}   {sum(buf[3:12])}")
