import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, raytrace_cpu_b200 as K
ctx=K.get_context(0)
for mode in (0,1,2):
    cnt=(C.c_int64*5)()
    K.lib.check(ctx.L.kt_selftest_divsqrt(ctx.h, mode, 4096, cnt))
    print('mode',mode, 'wrong div', cnt[0], 'wrong sqrt', cnt[1], 'slow div', cnt[2], 'slow sqrt', cnt[3], 'samples', cnt[4])
