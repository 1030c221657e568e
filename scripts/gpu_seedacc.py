import sys, ctypes as C, math
sys.path.insert(0,'/root/repo')
import raytrace_cpu_b200 as K
ctx=K.get_context(0)
cnt=(C.c_int64*5)()
K.lib.check(ctx.L.kt_selftest_divsqrt(ctx.h, 6, 2048, cnt))
for name,v in (("rcp64h",cnt[0]),("rsq64h (1 - x y^2)",cnt[1])):
    e=v/2.0**40
    print(f"max seed error {name}: {e:.3e} = 2^{math.log2(e):.2f}   ({cnt[4]} samples)")
for mode in (100, 101):
    cnt=(C.c_int64*5)()
    K.lib.check(ctx.L.kt_selftest_divsqrt(ctx.h, mode, 4096, cnt))
    print("lean rhs vs momenta(), spin", "+" if mode==100 else "-", "max rel diff:", {n: f"{v/2.0**60:.2e}" for n,v in zip(("pt","pphi","|ptheta|","|pr|","s2/c2"),cnt)})
