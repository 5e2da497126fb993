/* TEST INFRASTRUCTURE. Exercises the nine functions of the reference's host helper library through
 * whichever my_lib.h is first on the include path:
 *   oracle/Makefile builds it against /root/reference/src/my_lib.h  -> oracle/_ref/ref_my_lib (golden stdout)
 *   the product Makefile builds it against include/my_lib.h          -> apps/selftest_my_lib
 * tests/test_my_lib.py requires the two outputs to be byte-identical. */
#include <stdio.h>
#include <stdlib.h>
#include "my_lib.h"

int main(void)
{
	const int m = 3, k = 4, n = 5;
	float  Af[12], Bf[20], Cf[15];
	double Ad[12], Bd[20], Cd[15];
	unsigned s = 12345u;
	for (int i = 0; i < 12; i++) { s = s * 1664525u + 1013904223u; Af[i] = (float)((int)(s >> 8) % 2001 - 1000) / 317.0f; Ad[i] = (double)Af[i] * 1.000001; }
	for (int i = 0; i < 20; i++) { s = s * 1664525u + 1013904223u; Bf[i] = (float)((int)(s >> 8) % 2001 - 1000) / 113.0f; Bd[i] = (double)Bf[i] / 3.0; }
	fmatrixMul(Af, Bf, Cf, m, n, k);
	printf("fmatrixMul bits:"); for (int i = 0; i < 15; i++) { union { float f; unsigned u; } v; v.f = Cf[i]; printf(" %08x", v.u); } printf("\n");
	SmatrixMul(Af, Bf, Cf, m, n, k);
	printf("SmatrixMul bits:"); for (int i = 0; i < 15; i++) { union { float f; unsigned u; } v; v.f = Cf[i]; printf(" %08x", v.u); } printf("\n");
	dmatrixMul(Ad, Bd, Cd, m, n, k);
	printf("dmatrixMul bits:"); for (int i = 0; i < 15; i++) { union { double f; unsigned long long u; } v; v.f = Cd[i]; printf(" %016llx", v.u); } printf("\n");
	print_cloud(Cd, 5, 5);
	print_cloud(Cd, 5, 6);
	printScloud(Cf, 5, 4);
	printScloud(Cf, 5, 9);
	print_darray(Cd, 7);
	printSarray(Cf, 7);
	int iv[6] = { 3, -1, 4, 1, -5, 9 };
	print_iarray(iv, 6);
	printIarray(iv, 6);
	return 0;
}
