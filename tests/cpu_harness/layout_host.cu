/* TEST INFRASTRUCTURE — host-only checks of the index arithmetic the CUDA kernels share (sxs_dev.cuh): the sort key of
 * a pose and the tiled layout of the K3 -> K4 hand-off.  Compiled with nvcc, runs without a GPU.  Prints "ok" or the
 * first violated property. */
#include <set>
#include <vector>

#include "sxs_dev.cuh"

static int fail(const char *what, int L)
{
	printf("FAILED: %s (L = %d)\n", what, L);
	return 1;
}

int main()
{
	const int Ls[] = {1, 3, 15, 20, 30, 40};
	for (int L : Ls) {
		const int nb = L + 1, N = 2 * L + 1, NP = sxs_row_pad(N);
		if (NP % 8 != 0 || NP < N || NP >= N + 8) return fail("row padding", L);
		const unsigned long long per_cell = sxs_keys_per_cell(N);
		/* one cell in full: every (g1, g2, a2) */
		const sxs_pose_digits cell = {3, nb - 1, L / 2, 0, 0, 0};
		std::vector<unsigned long long> keys;
		for (int g1 = 0; g1 < N; g1++) {
			for (int g2 = 0; g2 < N; g2++) {
				for (int a2 = 0; a2 < N; a2 += (L > 20 ? 7 : 1)) {
					sxs_pose_digits d = cell;
					d.g1 = g1; d.g2 = g2; d.a2 = a2;
					const unsigned long long k = sxs_key_pack(nb, N, d);
					const sxs_pose_digits u = sxs_key_unpack(nb, N, k);
					if (u.z != d.z || u.b1 != d.b1 || u.b2 != d.b2 || u.g1 != g1 || u.g2 != g2 || u.a2 != a2) return fail("round trip", L);
					/* cells are contiguous key ranges, ordered by (z, b2, b1) */
					if (k / per_cell != ((unsigned long long)d.z * nb + d.b2) * nb + d.b1) return fail("cell prefix", L);
					/* points that differ only in a2 are neighbours */
					sxs_pose_digits e = d;
					e.a2 = (a2 + 1) % N;
					if (sxs_key_pack(nb, N, e) / N != k / N) return fail("a2 run", L);
					/* inside the cell the 128-byte band of the ligand operand is the most significant digit */
					if ((k % per_cell) / ((unsigned long long)N * 8 * N) != (unsigned long long)(g2 / 8)) return fail("band digit", L);
					keys.push_back(k);
				}
			}
		}
		if (std::set<unsigned long long>(keys.begin(), keys.end()).size() != keys.size()) return fail("keys distinct", L);
		/* neighbouring cells do not overlap, z is the most significant digit */
		sxs_pose_digits lo = {0, 0, 0, N - 1, N - 1, N - 1}, hi = {0, 1, 0, 0, 0, 0};
		if (!(sxs_key_pack(nb, N, lo) < sxs_key_pack(nb, N, hi))) return fail("cell order b1", L);
		lo = {0, nb - 1, nb - 1, N - 1, N - 1, N - 1}; hi = {1, 0, 0, 0, 0, 0};
		if (!(sxs_key_pack(nb, N, lo) < sxs_key_pack(nb, N, hi))) return fail("cell order z", L);
		lo = {0, nb - 1, 0, N - 1, N - 1, N - 1}; hi = {0, 0, 1, 0, 0, 0};
		if (!(sxs_key_pack(nb, N, lo) < sxs_key_pack(nb, N, hi))) return fail("cell order b2 over b1", L);
		/* the largest key of a 4096-step z table fits 64 bits with room for the all-ones sentinel */
		hi = {4095, nb - 1, nb - 1, N - 1, N - 1, N - 1};
		if (sxs_key_pack(nb, N, hi) >= 0xFFFFFFFFFFFFFFFFull / 2) return fail("key range", L);
	}
	/* hand-off layout: a bijection from (point, q, k) onto [0, npts * qnum * 6); the six terms of a node are contiguous */
	for (int qnum : {1, 50, 100}) {
		const long long npts = 32 * 5 + 7, tiles = (npts + 31) / 32;
		std::vector<char> seen((size_t)tiles * qnum * 6 * 32, 0);
		for (long long p = 0; p < npts; p++) {
			for (int q = 0; q < qnum; q++) {
				for (int k = 0; k < 6; k++) {
					const size_t i = sxs_x_index(p, qnum, q, k);
					if (i >= seen.size() || seen[i]) return fail("x index bijection", qnum);
					seen[i] = 1;
					if (k < 5 && sxs_x_index(p, qnum, q, k + 1) != i + 1) return fail("x index term stride", qnum);
					if (i % 2 == 0 && k % 2 != 0) return fail("x index 16-byte pairs", qnum);
				}
			}
		}
	}
	printf("ok\n");
	return 0;
}
