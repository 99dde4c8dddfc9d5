// Host check of dwt_b200/csrc/bitslice.cuh (the 16 x 16 bit-matrix transpose behind the TMA Hilbert kernels):
// plane word p of 32 coefficients must hold bit p of every |coefficient| (bit 15: the sign), and the transpose
// must be its own inverse.  Built and run by tests/test_bitslice.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "bitslice.cuh"

int main(int argc, char **argv)
{
	const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
	uint32_t seed = 12345u;
	auto rnd = [&]() {
		seed = seed * 1664525u + 1013904223u;
		return seed >> 8;
	};
	for (int r = 0; r < rounds; ++r) {
		int c[32];
		const int bits = 1 + (int)(rnd() % 15); // magnitudes below 2^bits: every plane count the kernels accept
		for (int i = 0; i < 32; ++i) {
			const int mag = (int)(rnd() & ((1u << bits) - 1u));
			c[i] = (rnd() & 1u) ? -mag : mag;
			if (r == 0)
				c[i] = i % 3 == 0 ? 0 : (i & 1 ? -32767 : 32767); // zeros and the extremes
		}
		uint32_t w[16], keep[16];
		for (int i = 0; i < 16; ++i)
			keep[i] = w[i] = bitslice_half(c[i]) | (bitslice_half(c[i + 16]) << 16);
		bitslice_transpose16(w);
		for (int p = 0; p < 16; ++p)
			for (int i = 0; i < 32; ++i) {
				const uint32_t mag = (uint32_t)(c[i] < 0 ? -c[i] : c[i]);
				const uint32_t want = p < 15 ? (mag >> p) & 1u : (c[i] < 0 ? 1u : 0u);
				if (((w[p] >> i) & 1u) != want) {
					printf("plane %d coefficient %d: round %d\n", p, i, r);
					return 1;
				}
			}
		bitslice_transpose16(w);
		for (int i = 0; i < 16; ++i) {
			if (w[i] != keep[i]) {
				printf("not an involution: round %d\n", r);
				return 1;
			}
			if (bitslice_value(keep[i] & 0xffffu) != c[i] || bitslice_value(keep[i] >> 16) != c[i + 16]) {
				printf("half-word round trip: round %d\n", r);
				return 1;
			}
		}
	}
	printf("ok %d\n", rounds);
	return 0;
}
