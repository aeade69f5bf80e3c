/* Test helper (tests/test_lowering.py): the shadow march's division pre-test as lol_kernel.cuh writes it
 * (LOL_DIV_PRETEST).  Whenever `num > RN(thr * t) && RN(thr * t) >= 2^-120` with
 * thr = RN(res * (1 + 2^-21)) (or inf for res < 2^-100), the reference's res = minf(res, num / t)
 * (float.h:6-13: MINSS, `res < q ? res : q`) must leave res unchanged, bit for bit.  Operands: random
 * bits, and quotients within a few thousand ulps of res, where the margin is decided. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s = 88172645463325252ull;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t ubits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
	const long N = argc > 1 ? atol(argv[1]) : 20000000;
	const float C = bits(0x3f800004u);
	long skipped = 0, bad = 0, near_skipped = 0;
	for (long i = 0; i < N; i++) {
		const uint64_t r = rnd(), r2 = rnd();
		float res, t, num;
		const int mode = (int)(i % 4);
		if (mode == 0) { /* anything */
			res = bits((uint32_t)r); t = bits((uint32_t)(r >> 32)); num = bits((uint32_t)r2);
		} else if (mode == 1) { /* plausible magnitudes */
			res = (float)((double)(r & 0xffffff) / 16777216.0); t = (float)((double)((r >> 24) & 0xffffff) / 1e5);
			num = (float)((double)(r2 & 0xffffff) / 1e4 - 200.0);
		} else { /* quotient close to res: num = res * t * (1 + k * 2^-24) */
			res = mode == 2 ? (float)((double)((r & 0xffffff) + 1) / 16777216.0) : bits(((uint32_t)r & 0x007fffffu) | (((uint32_t)(r >> 23) % 120u + 8u) << 23));
			t = bits(((uint32_t)(r >> 32) & 0x007fffffu) | (((uint32_t)(r >> 55) % 60u + 97u) << 23));
			num = (float)((double)res * (double)t * (1.0 + (double)((long)(r2 % 64) - 16) * 0x1p-24));
		}
		if (!(res > 0.f)) /* the loop only continues with res > 0 (or NaN: handled below) */
			res = bits(0x7fc00000u);
		{
			const float thr = (res >= 0x1p-100f) ? res * C : INFINITY;
			const float bound = thr * t;
			const float q = num / t;
			const float want = (res < q) ? res : q;
			if (num > bound && bound >= 0x1p-120f) {
				skipped++;
				near_skipped += mode >= 2;
				if (ubits(want) != ubits(res)) {
					if (bad++ < 10)
						printf("res=%a t=%a num=%a: q=%a, minf gives %a\n", res, t, num, q, want);
				}
			}
		}
	}
	printf("%ld cases, %ld skipped the division (%ld of them near the margin), %ld mismatches\n", N, skipped, near_skipped, bad);
	return bad != 0 || skipped == 0 || near_skipped == 0;
}
