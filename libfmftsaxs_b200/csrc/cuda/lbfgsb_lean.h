/* Register-resident form of the per-fit optimiser: L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; Morales & Nocedal 2011)
 * for n = 2 boxed variables and m = 3 corrections, as the reference drives it (src/min_saxs.c:196-259 over the
 * vendored lbfgsb/src/).  It computes, operation for operation, what lbfgsb_n2m3.h computes (the restatement that
 * is pinned bit for bit against the vendored code on the host); what differs is where the state lives:
 *
 *   - the three correction pairs are kept in LOGICAL order (column 1 = oldest) and shifted when the memory is full,
 *     instead of the circular head/tail pointers of the vendored code — the values and every sum over them are the
 *     same, but no index depends on run-time state any more;
 *   - the 2m x 2m matrix of the subspace step keeps its two diagonal blocks at the fixed offset m (the vendored code
 *     packs them at offset col), the 2col vectors are held as two halves;
 *   - the free / entering / leaving variable lists over two variables are flags; sums over such a list have at most
 *     two terms, and IEEE addition of two terms does not depend on their order (0 + a + b == 0 + b + a);
 *   - every loop has a compile-time trip count (3, 2) with a guard on col, so that after unrolling every array
 *     element is a scalar the compiler can keep in a register: the optimiser touches no memory at all.
 *
 * K4 before this header: ~1.1 KB of thread-local state per fit, 230 GB of DRAM traffic per 4.44 M fits from that state
 * alone, ~21 000 executed instructions per iteration boundary (index arithmetic, rolled loops, local loads).
 *
 * Third-party notice: L-BFGS-B 3.0 (C. Zhu, R. Byrd, J. Nocedal, J. L. Morales), C version "L-BFGS-B-C"
 * (c) 2015 Stephen Becker, BSD 3-clause; the full notice is carried in lbfgsb_n2m3.h, which this file accompanies.
 */
#ifndef SXS_LBFGSB_LEAN_H
#define SXS_LBFGSB_LEAN_H

#include <math.h>

#ifndef SXS_HD
#ifdef __CUDACC__
#define SXS_HD __host__ __device__ __forceinline__
#else
#define SXS_HD static inline
#endif
#endif

#ifdef __CUDACC__
#define LQ_FN __host__ __device__ __forceinline__
#define LQ_UNROLL _Pragma("unroll")
#define LQ_ROLLED _Pragma("unroll 1")
#else
#define LQ_FN static inline __attribute__((always_inline))
#define LQ_UNROLL _Pragma("GCC unroll 8")
#define LQ_ROLLED
#endif

#define LQ_M 3
/* lbfgsb/src/lbfgsb.h:213-222 */
#define LQ_FTOL 1.0e-3
#define LQ_GTOL 0.9
#define LQ_XTOL 0.1
#define LQ_STPMIN 0.0
#define LQ_EPSMCH 2.220446049250313e-16

/* the box (src/define.h:28-34) is a compile-time constant of the fit */
#ifndef LQ_L1
#define LQ_L1 0.96
#define LQ_U1 1.04
#define LQ_L2 (-2.00)
#define LQ_U2 4.00
#endif
#define LQ_LO(i) ((i) == 1 ? LQ_L1 : LQ_L2)
#define LQ_UP(i) ((i) == 1 ? LQ_U1 : LQ_U2)

enum lq_search_task { LQ_LS_START = 0, LQ_LS_FG, LQ_LS_CONVERGENCE, LQ_LS_WARNING };
enum lq_phase { LQ_PH_INIT = 0, LQ_PH_FIRST_EVAL, LQ_PH_LINESEARCH, LQ_PH_DONE,
                LQ_PH_B_FIRST, LQ_PH_B_ACCEPTED, LQ_PH_B_NEW_ITER };
enum lq_status { LQ_DONE = 0, LQ_NEED_EVAL = 1, LQ_NEED_B = 2 };

/* 1-based like the algorithm papers; element 0 of every array is never touched (and costs nothing once the struct
 * is scalarised). */
struct lq_state {
	double x[3], g[3];
	double f;
	double ws[3][4], wy[3][4];        /* [variable][correction], correction 1 = oldest */
	double sy[4][4], ss[4][4], wt[4][4];
	double sq[4];                     /* sqrt(sy[i][i]): bmv takes these roots six times per call, sy changes once per iteration */
	double wn1[7][7];                 /* lower triangle of N: rows/cols 1..3 = Y block, 4..6 = S block */
	double wn[7][7];                  /* upper-triangular factor: blocks at 1..col and LQ_M+1..LQ_M+col */
	double z[3], r[3], d[3], t[3];
	double c1h[4], c2h[4];            /* Cauchy-point coefficients c = W'(xcp - x): Y half, S half */
	double theta, fold, gd, gdold, stp, stpmx, sbgnrm, dtd;
	int col, iupdat, updatd, iback, ifun, iter, nfgv;
	int free1, free2;                 /* variable i is free at the Cauchy point (index[] of the vendored code) */
	int iw1, iw2;                     /* iwhere[] */
	int ent1, ent2, lv1, lv2;         /* entering / leaving the free set this iteration (indx2[]) */
	int wrk;
	int phase;
	int brackt, stage, ls_task;
	double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

/* IEEE division and square root.  On the device they are real calls: the optimiser has ~150 division sites and
 * nvcc expands each into ~25 instructions (reciprocal seed, Newton steps, slow-path test); inlined they are half of
 * the kernel's code (11 300 SASS instructions, beyond the instruction cache), as calls they cost ~5. */
#if defined(__CUDA_ARCH__) && !defined(LQ_INLINE_DIV)
static __device__ __noinline__ double lq_div(double a, double b)
{
	/* 0 / b: a quarter of the optimiser's quotients while the memory fills (zero columns of the correction matrices);
	 * nvcc's division sends a zero numerator down its slow path (r2n ncu: 2 % of the kernel's samples at 7.5 lanes) */
	if (a == 0.0 && b == b && b != 0.0) {
		return __longlong_as_double((__double_as_longlong(a) ^ __double_as_longlong(b)) & (long long)0x8000000000000000ull);
	}
	return __ddiv_rn(a, b);
}
static __device__ __noinline__ double lq_sqrt(double a) { return __dsqrt_rn(a); }
#else
LQ_FN double lq_div(double a, double b) { return a / b; }
LQ_FN double lq_sqrt(double a) { return sqrt(a); }
#endif

LQ_FN double lq_abs(double a) { return a >= 0 ? a : -a; }
LQ_FN double lq_max(double a, double b) { return a >= b ? a : b; }
LQ_FN double lq_min(double a, double b) { return a <= b ? a : b; }

LQ_FN void lq_reset_memory(struct lq_state *s)
{
	s->col = 0;
	s->theta = 1.0;
	s->iupdat = 0;
	s->updatd = 0;
}

/* LINPACK dpofa on the n x n block of a[7][7] (or a[4][4]) whose (1,1) element is a[off+1][off+1]; n <= 3.
 * A is a macro argument so that the same text serves both array shapes with constant indices. */
#define LQ_DPOFA(A, off, n, info)                                              \
	do {                                                                       \
		(info) = 0;                                                            \
		LQ_UNROLL                                                              \
		for (int j_ = 1; j_ <= LQ_M; j_++) {                                   \
			if (j_ <= (n) && (info) == 0) {                                    \
				double sacc_ = 0.0;                                            \
				LQ_UNROLL                                                      \
				for (int k_ = 1; k_ <= LQ_M - 1; k_++) {                       \
					if (k_ <= j_ - 1) {                                        \
						double dot_ = 0.0;                                     \
						LQ_UNROLL                                              \
						for (int i_ = 1; i_ <= LQ_M - 2; i_++) {               \
							if (i_ <= k_ - 1) {                                \
								dot_ += A[(off) + i_][(off) + k_] * A[(off) + i_][(off) + j_]; \
							}                                                  \
						}                                                      \
						double tt_ = A[(off) + k_][(off) + j_] - dot_;         \
						tt_ = lq_div(tt_, A[(off) + k_][(off) + k_]);          \
						A[(off) + k_][(off) + j_] = tt_;                       \
						sacc_ += tt_ * tt_;                                    \
					}                                                          \
				}                                                              \
				sacc_ = A[(off) + j_][(off) + j_] - sacc_;                     \
				if (sacc_ <= 0.0) {                                            \
					(info) = j_;                                               \
				} else {                                                       \
					A[(off) + j_][(off) + j_] = lq_sqrt(sacc_);                \
				}                                                              \
			}                                                                  \
		}                                                                      \
	} while (0)

/* LINPACK dtrsl with the col x col upper-triangular factor wt: job 11 solves trans(T) x = b, job 01 solves T x = b;
 * b[1..col].  Returns the index of a zero diagonal element or 0. */
LQ_FN int lq_dtrsl_wt(const struct lq_state *s, double *b, const int transposed)
{
	const int n = s->col;
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= n && s->wt[i][i] == 0.0) {
			return i;
		}
	}
	if (transposed) {
		b[1] = lq_div(b[1], s->wt[1][1]);
		LQ_UNROLL
		for (int j = 2; j <= LQ_M; j++) {
			if (j <= n) {
				double dot = 0.0;
				LQ_UNROLL
				for (int i = 1; i <= LQ_M - 1; i++) {
					if (i <= j - 1) {
						dot += s->wt[i][j] * b[i];
					}
				}
				b[j] -= dot;
				b[j] = lq_div(b[j], s->wt[j][j]);
			}
		}
	} else {
		/* b[n] /= t[n][n], then jj = 2..n with j = n - jj + 1 */
		LQ_UNROLL
		for (int j = LQ_M; j >= 1; j--) {
			if (j == n) {
				b[j] = lq_div(b[j], s->wt[j][j]);
			} else if (j < n) {
				const double temp = -b[j + 1];
				if (temp != 0.0) {
					LQ_UNROLL
					for (int i = 1; i <= LQ_M - 1; i++) {
						if (i <= j) {
							b[i] += temp * s->wt[i][j + 1];
						}
					}
				}
				b[j] = lq_div(b[j], s->wt[j][j]);
			}
		}
	}
	return 0;
}

/* Product of the 2m x 2m middle matrix with v = (v1, v2) (subalgorithms.c bmv, :120-259); halves are [1..col]. */
LQ_FN int lq_bmv(const struct lq_state *s, const double *v1, const double *v2, double *p1, double *p2)
{
	const int col = s->col;
	if (col == 0) {
		return 0;
	}
	p2[1] = v2[1];
	LQ_UNROLL
	for (int i = 2; i <= LQ_M; i++) {
		if (i <= col) {
			double sum = 0.0;
			LQ_UNROLL
			for (int k = 1; k <= LQ_M - 1; k++) {
				if (k <= i - 1) {
					sum += lq_div(s->sy[i][k] * v1[k], s->sy[k][k]);
				}
			}
			p2[i] = v2[i] + sum;
		}
	}
	int info = lq_dtrsl_wt(s, p2, 1);
	if (info != 0) {
		return info;
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) {
			p1[i] = lq_div(v1[i], s->sq[i]);
		}
	}
	info = lq_dtrsl_wt(s, p2, 0);
	if (info != 0) {
		return info;
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) {
			p1[i] = lq_div(-p1[i], s->sq[i]);
		}
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) {
			double sum = 0.0;
			LQ_UNROLL
			for (int k = 2; k <= LQ_M; k++) {
				if (k >= i + 1 && k <= col) {
					sum += lq_div(s->sy[k][i] * p2[k], s->sy[i][i]);
				}
			}
			p1[i] += sum;
		}
	}
	return 0;
}

/* subalgorithms.c projgr, :1513-1560 */
LQ_FN double lq_projgr(const struct lq_state *s)
{
	double sbgnrm = 0.0;
	LQ_UNROLL
	for (int i = 1; i <= 2; i++) {
		double gi = s->g[i];
		if (gi < 0.0) {
			gi = lq_max(s->x[i] - LQ_UP(i), gi);
		} else {
			gi = lq_min(s->x[i] - LQ_LO(i), gi);
		}
		sbgnrm = lq_max(sbgnrm, lq_abs(gi));
	}
	return sbgnrm;
}

LQ_FN int lq_iw(const struct lq_state *s, int i) { return i == 1 ? s->iw1 : s->iw2; }
LQ_FN void lq_set_iw(struct lq_state *s, int i, int v) { if (i == 1) s->iw1 = v; else s->iw2 = v; }

/* Generalised Cauchy point (subalgorithms.c cauchy, :261-818).  p, wbp, v are scratch; c is kept for cmprlb. */
LQ_FN int lq_cauchy(struct lq_state *s)
{
	double p1[4], p2[4], wb1[4], wb2[4], v1[4], v2[4];
	double tbk1 = 0.0, tbk2 = 0.0;   /* breakpoint times */
	int ibk1 = 0, ibk2 = 0;          /* their variables */
	double *d = s->d, *xcp = s->z;
	const int col = s->col;
	const double theta = s->theta;

	/* c is zeroed here rather than after the two early exits below: both need a vanishing projected gradient, with
	 * which mainlb has already stopped (pgtol = 1e-5 > 0), so no path reads c between them and this point — and c
	 * need not survive from one iteration to the next */
	LQ_UNROLL
	for (int j = 1; j <= LQ_M; j++) {
		s->c1h[j] = 0.0; s->c2h[j] = 0.0;
	}
	if (s->sbgnrm <= 0.0) {
		xcp[1] = s->x[1];
		xcp[2] = s->x[2];
		return 0;
	}
	int bnded = 1;
	int nzero = 0; /* variables with a zero gradient component (the vendored code's nfree counter) */
	int nbreak = 0;
	int ibkmin = 0;
	double bkmin = 0.0;
	double f1 = 0.0;
	double tl = 0.0, tu = 0.0;

	LQ_UNROLL
	for (int j = 1; j <= LQ_M; j++) {
		p1[j] = 0.0; p2[j] = 0.0;
	}
	LQ_UNROLL
	for (int i = 1; i <= 2; i++) {
		const double neggi = -s->g[i];
		int iw = lq_iw(s, i);
		if (iw != 3 && iw != -1) {
			tl = s->x[i] - LQ_LO(i);
			tu = LQ_UP(i) - s->x[i];
			const int xlower = tl <= 0.0;
			const int xupper = tu <= 0.0;
			iw = 0;
			if (xlower) {
				if (neggi <= 0.0) {
					iw = 1;
				}
			} else if (xupper) {
				if (neggi >= 0.0) {
					iw = 2;
				}
			} else {
				if (lq_abs(neggi) <= 0.0) {
					iw = -3;
				}
			}
			lq_set_iw(s, i, iw);
		}
		if (iw != 0 && iw != -1) {
			d[i] = 0.0;
		} else {
			d[i] = neggi;
			f1 -= neggi * neggi;
			LQ_UNROLL
			for (int j = 1; j <= LQ_M; j++) {
				if (j <= col) {
					p1[j] += s->wy[i][j] * neggi;
					p2[j] += s->ws[i][j] * neggi;
				}
			}
			if (neggi < 0.0) {
				++nbreak;
				const double tb = lq_div(tl, -neggi);
				if (nbreak == 1) { tbk1 = tb; ibk1 = i; } else { tbk2 = tb; ibk2 = i; }
				if (nbreak == 1 || tb < bkmin) {
					bkmin = tb;
					ibkmin = nbreak;
				}
			} else if (neggi > 0.0) {
				++nbreak;
				const double tb = lq_div(tu, neggi);
				if (nbreak == 1) { tbk1 = tb; ibk1 = i; } else { tbk2 = tb; ibk2 = i; }
				if (nbreak == 1 || tb < bkmin) {
					bkmin = tb;
					ibkmin = nbreak;
				}
			} else {
				++nzero;
				if (lq_abs(neggi) > 0.0) {
					bnded = 0;
				}
			}
		}
	}
	if (theta != 1.0) {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) {
				p2[j] = theta * p2[j];
			}
		}
	}
	xcp[1] = s->x[1];
	xcp[2] = s->x[2];
	if (nbreak == 0 && nzero == 0) {
		return 0;
	}
	double f2 = -theta * f1;
	const double f2_org = f2;
	if (col > 0) {
		const int info = lq_bmv(s, p1, p2, v1, v2);
		if (info != 0) {
			return info;
		}
		double dot = 0.0;
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) dot += v1[j] * p1[j];
		}
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) dot += v2[j] * p2[j];
		}
		f2 -= dot;
	}
	double dtm = lq_div(-f1, f2);
	double tsum = 0.0;
	int skip_to_end = 0;

	if (nbreak != 0) {
		int nleft = nbreak;
		double tj = 0.0;
		/* at most two breakpoints: the smaller one first, then the other (kept as a loop: one copy of the code) */
		LQ_ROLLED
		for (int iter = 1; iter <= 2; iter++) {
			if (iter > nbreak || skip_to_end == 2) {
				break;
			}
			const double tj0 = tj;
			int ibp;
			if (iter == 1) {
				tj = bkmin;
				ibp = ibkmin == 1 ? ibk1 : ibk2;
			} else {
				/* the remaining breakpoint (the heap of the vendored code has one element left) */
				tj = ibkmin == 1 ? tbk2 : tbk1;
				ibp = ibkmin == 1 ? ibk2 : ibk1;
			}
			const double dt = tj - tj0;
			if (dtm < dt) {
				skip_to_end = 2; /* leave the loop, finish normally */
				break;
			}
			tsum += dt;
			--nleft;
			const double dibp = ibp == 1 ? d[1] : d[2];
			if (ibp == 1) d[1] = 0.0; else d[2] = 0.0;
			double zibp;
			const double xi = ibp == 1 ? s->x[1] : s->x[2];
			const double ui = ibp == 1 ? LQ_U1 : LQ_U2, li = ibp == 1 ? LQ_L1 : LQ_L2;
			if (dibp > 0.0) {
				zibp = ui - xi;
				if (ibp == 1) xcp[1] = ui; else xcp[2] = ui;
				lq_set_iw(s, ibp, 2);
			} else {
				zibp = li - xi;
				if (ibp == 1) xcp[1] = li; else xcp[2] = li;
				lq_set_iw(s, ibp, 1);
			}
			if (nleft == 0 && nbreak == 2) {
				dtm = dt;
				skip_to_end = 1;
				break;
			}
			const double dibp2 = dibp * dibp;
			f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
			f2 -= theta * dibp2;
			if (col > 0) {
				if (dt != 0.0) {
					LQ_UNROLL
					for (int j = 1; j <= LQ_M; j++) {
						if (j <= col) s->c1h[j] += dt * p1[j];
					}
					LQ_UNROLL
					for (int j = 1; j <= LQ_M; j++) {
						if (j <= col) s->c2h[j] += dt * p2[j];
					}
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) {
						wb1[j] = ibp == 1 ? s->wy[1][j] : s->wy[2][j];
						wb2[j] = theta * (ibp == 1 ? s->ws[1][j] : s->ws[2][j]);
					}
				}
				const int info = lq_bmv(s, wb1, wb2, v1, v2);
				if (info != 0) {
					return info;
				}
				double wmc = 0.0, wmp = 0.0, wmw = 0.0;
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmc += s->c1h[j] * v1[j];
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmc += s->c2h[j] * v2[j];
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmp += p1[j] * v1[j];
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmp += p2[j] * v2[j];
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmw += wb1[j] * v1[j];
				}
				LQ_UNROLL
				for (int j = 1; j <= LQ_M; j++) {
					if (j <= col) wmw += wb2[j] * v2[j];
				}
				const double mdibp = -dibp;
				if (mdibp != 0.0) {
					LQ_UNROLL
					for (int j = 1; j <= LQ_M; j++) {
						if (j <= col) p1[j] += mdibp * wb1[j];
					}
					LQ_UNROLL
					for (int j = 1; j <= LQ_M; j++) {
						if (j <= col) p2[j] += mdibp * wb2[j];
					}
				}
				f1 += dibp * wmc;
				f2 = f2 + dibp * 2.0 * wmp - dibp2 * wmw;
			}
			f2 = lq_max(LQ_EPSMCH * f2_org, f2);
			if (nleft > 0) {
				dtm = lq_div(-f1, f2);
				continue;
			} else if (bnded) {
				f1 = 0.0;
				f2 = 0.0;
				dtm = 0.0;
			} else {
				dtm = lq_div(-f1, f2);
			}
			break;
		}
	}
	if (skip_to_end != 1) {
		if (dtm <= 0.0) {
			dtm = 0.0;
		}
		tsum += dtm;
		if (tsum != 0.0) {
			xcp[1] += tsum * d[1];
			xcp[2] += tsum * d[2];
		}
	}
	if (col > 0 && dtm != 0.0) {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) s->c1h[j] += dtm * p1[j];
		}
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) s->c2h[j] += dtm * p2[j];
		}
	}
	return 0;
}

/* Free/active bookkeeping at the Cauchy point (subalgorithms.c freev, :1741-1814). */
LQ_FN void lq_freev(struct lq_state *s)
{
	s->ent1 = s->ent2 = s->lv1 = s->lv2 = 0;
	if (s->iter > 0) {
		/* previously free and now at a bound: leaves; previously bound and now free: enters */
		if (s->free1 && s->iw1 > 0) s->lv1 = 1;
		if (s->free2 && s->iw2 > 0) s->lv2 = 1;
		if (!s->free1 && s->iw1 <= 0) s->ent1 = 1;
		if (!s->free2 && s->iw2 <= 0) s->ent2 = 1;
	}
	s->wrk = (s->lv1 | s->lv2) || (s->ent1 | s->ent2) || s->updatd;
	s->free1 = s->iw1 <= 0;
	s->free2 = s->iw2 <= 0;
}

/* inner products over a set of the two variables given by flags: at most two terms, order-free */
#define LQ_DOT2(f1, f2, A, ia, B, ib) (((f1) ? A[1][ia] * B[1][ib] : 0.0) + ((f2) ? A[2][ia] * B[2][ib] : 0.0))

/* sum over a flagged set exactly as a loop `t = 0; for k in set: t += term_k` computes it */
LQ_FN double lq_setsum(int f1, double t1, int f2, double t2)
{
	double t = 0.0;
	if (f1) t += t1;
	if (f2) t += t2;
	return t;
}

/* LEL^T factorisation of the indefinite subspace matrix (subalgorithms.c formk, :820-1303). */
LQ_FN int lq_formk(struct lq_state *s)
{
	const int col = s->col;
	const int fr1 = s->free1, fr2 = s->free2;      /* Z: free variables */
	const int ac1 = !s->free1, ac2 = !s->free2;    /* A: active variables */
	int upcl;

	/* The vendored code skips formk altogether when neither the free set nor the memory changed (wrk == 0) and
	 * reuses the factor from the iteration before.  Here only the update of wn1 is skipped: wn1, theta, sy and col
	 * are then unchanged too, so rebuilding and refactorising wn yields the very same factor, and wn need not
	 * survive from one iteration to the next (21 doubles less to keep per fit). */
	if (s->wrk) {
	if (s->updatd) {
		if (s->iupdat > LQ_M) {
			/* shift old part of WN1 */
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M - 1; jy++) {
				const int js = LQ_M + jy;
				LQ_UNROLL
				for (int i = 0; i < LQ_M - 1; i++) {
					if (i < LQ_M - jy) s->wn1[jy + i][jy] = s->wn1[jy + 1 + i][jy + 1];
				}
				LQ_UNROLL
				for (int i = 0; i < LQ_M - 1; i++) {
					if (i < LQ_M - jy) s->wn1[js + i][js] = s->wn1[js + 1 + i][js + 1];
				}
				LQ_UNROLL
				for (int i = 0; i < LQ_M - 1; i++) {
					s->wn1[LQ_M + 1 + i][jy] = s->wn1[LQ_M + 2 + i][jy + 1];
				}
			}
		}
		/* put new rows in blocks (1,1), (2,1) and (2,2): row col against columns jy = 1..col (newest pair = col) */
		LQ_UNROLL
		for (int cc = 1; cc <= LQ_M; cc++) {
			if (cc == col) {
				LQ_UNROLL
				for (int jy = 1; jy <= LQ_M; jy++) {
					if (jy <= cc) {
						const int js = LQ_M + jy;
						const double temp1 = lq_setsum(fr1, s->wy[1][cc] * s->wy[1][jy], fr2, s->wy[2][cc] * s->wy[2][jy]);
						const double temp2 = lq_setsum(ac1, s->ws[1][cc] * s->ws[1][jy], ac2, s->ws[2][cc] * s->ws[2][jy]);
						const double temp3 = lq_setsum(ac1, s->ws[1][cc] * s->wy[1][jy], ac2, s->ws[2][cc] * s->wy[2][jy]);
						s->wn1[cc][jy] = temp1;
						s->wn1[LQ_M + cc][js] = temp2;
						s->wn1[LQ_M + cc][jy] = temp3;
					}
				}
				/* put new column in block (2,1) */
				LQ_UNROLL
				for (int i = 1; i <= LQ_M; i++) {
					if (i <= cc) {
						const double temp3 = lq_setsum(fr1, s->ws[1][i] * s->wy[1][cc], fr2, s->ws[2][i] * s->wy[2][cc]);
						s->wn1[LQ_M + i][cc] = temp3;
					}
				}
			}
		}
		upcl = col - 1;
	} else {
		upcl = col;
	}
	/* modify the old parts in blocks (1,1) and (2,2) due to changes in the set of free variables */
	LQ_UNROLL
	for (int iy = 1; iy <= LQ_M; iy++) {
		if (iy <= upcl) {
			const int is = LQ_M + iy;
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M; jy++) {
				if (jy <= iy) {
					const int js = LQ_M + jy;
					const double temp1 = lq_setsum(s->ent1, s->wy[1][iy] * s->wy[1][jy], s->ent2, s->wy[2][iy] * s->wy[2][jy]);
					const double temp2 = lq_setsum(s->ent1, s->ws[1][iy] * s->ws[1][jy], s->ent2, s->ws[2][iy] * s->ws[2][jy]);
					const double temp3 = lq_setsum(s->lv1, s->wy[1][iy] * s->wy[1][jy], s->lv2, s->wy[2][iy] * s->wy[2][jy]);
					const double temp4 = lq_setsum(s->lv1, s->ws[1][iy] * s->ws[1][jy], s->lv2, s->ws[2][iy] * s->ws[2][jy]);
					s->wn1[iy][jy] = s->wn1[iy][jy] + temp1 - temp3;
					s->wn1[is][js] = s->wn1[is][js] - temp2 + temp4;
				}
			}
		}
	}
	/* modify the old parts in block (2,1) */
	LQ_UNROLL
	for (int ii = 1; ii <= LQ_M; ii++) {
		if (ii <= upcl) {
			const int is = LQ_M + ii;
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M; jy++) {
				if (jy <= upcl) {
					const double temp1 = lq_setsum(s->ent1, s->ws[1][ii] * s->wy[1][jy], s->ent2, s->ws[2][ii] * s->wy[2][jy]);
					const double temp3 = lq_setsum(s->lv1, s->ws[1][ii] * s->wy[1][jy], s->lv2, s->ws[2][ii] * s->wy[2][jy]);
					if (is <= jy + LQ_M) {
						s->wn1[is][jy] = s->wn1[is][jy] + temp1 - temp3;
					} else {
						s->wn1[is][jy] = s->wn1[is][jy] - temp1 + temp3;
					}
				}
			}
		}
	}
	} /* wrk */
	/* form the upper triangle of WN = [D+Y'ZZ'Y/theta   -L_a'+R_z' ; -L_a+R_z   S'AA'S*theta]; the S block sits at
	 * the fixed offset LQ_M here (the vendored code packs it at offset col) */
	const double theta = s->theta;
	LQ_UNROLL
	for (int iy = 1; iy <= LQ_M; iy++) {
		if (iy <= col) {
			const int is = LQ_M + iy;
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M; jy++) {
				if (jy <= iy) {
					const int js = LQ_M + jy;
					s->wn[jy][iy] = lq_div(s->wn1[iy][jy], theta);
					s->wn[js][is] = s->wn1[is][js] * theta;
				}
			}
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M; jy++) {
				if (jy <= iy - 1) {
					s->wn[jy][is] = -s->wn1[is][jy];
				}
			}
			LQ_UNROLL
			for (int jy = 1; jy <= LQ_M; jy++) {
				if (jy >= iy && jy <= col) {
					s->wn[jy][is] = s->wn1[is][jy];
				}
			}
			s->wn[iy][iy] += s->sy[iy][iy];
		}
	}
	/* first Cholesky: (1,1) block of WN */
	int info;
	LQ_DPOFA(s->wn, 0, col, info);
	if (info != 0) {
		return -1;
	}
	/* then form L^-1(-L_a'+R_z') in the (1,2) block: each column js, rows 1..col, solved against trans(T11) */
	LQ_UNROLL
	for (int jj = 1; jj <= LQ_M; jj++) {
		if (jj <= col) {
			const int js = LQ_M + jj;
			/* dtrsl job 11 on the leading block of wn (its diagonal is non-zero after a successful dpofa unless
			 * an element underflowed; the vendored code ignores info here as well) */
			int zero = 0;
			LQ_UNROLL
			for (int i = 1; i <= LQ_M; i++) {
				if (i <= col && s->wn[i][i] == 0.0) zero = 1;
			}
			if (!zero) {
				s->wn[1][js] = lq_div(s->wn[1][js], s->wn[1][1]);
				LQ_UNROLL
				for (int j = 2; j <= LQ_M; j++) {
					if (j <= col) {
						double dot = 0.0;
						LQ_UNROLL
						for (int i = 1; i <= LQ_M - 1; i++) {
							if (i <= j - 1) dot += s->wn[i][j] * s->wn[i][js];
						}
						s->wn[j][js] -= dot;
						s->wn[j][js] = lq_div(s->wn[j][js], s->wn[j][j]);
					}
				}
			}
		}
	}
	/* form S'AA'S*theta + (L^-1(-L_a'+R_z'))'(L^-1(-L_a'+R_z')) in the upper triangle of (2,2) */
	LQ_UNROLL
	for (int ii = 1; ii <= LQ_M; ii++) {
		if (ii <= col) {
			const int is = LQ_M + ii;
			LQ_UNROLL
			for (int jj = 1; jj <= LQ_M; jj++) {
				if (jj >= ii && jj <= col) {
					const int js = LQ_M + jj;
					double dot = 0.0;
					LQ_UNROLL
					for (int i = 1; i <= LQ_M; i++) {
						if (i <= col) dot += s->wn[i][is] * s->wn[i][js];
					}
					s->wn[is][js] += dot;
				}
			}
		}
	}
	/* Cholesky factorisation of the (2,2) block */
	LQ_DPOFA(s->wn, LQ_M, col, info);
	if (info != 0) {
		return -2;
	}
	return 0;
}

/* r = -Z'B(xcp - xk) - Z'g (subalgorithms.c cmprlb, :1305-1391); r is indexed by position in the free list. */
LQ_FN int lq_cmprlb(struct lq_state *s)
{
	const int col = s->col;
	double p1[4], p2[4];
	/* free list: (1,2) if both free, else the single free variable at position 1 */
	const int k1 = s->free1 ? 1 : 2;          /* variable at position 1 (if nfree >= 1) */
	const int nfree = s->free1 + s->free2;
	if (nfree >= 1) {
		s->r[1] = k1 == 1 ? -s->theta * (s->z[1] - s->x[1]) - s->g[1] : -s->theta * (s->z[2] - s->x[2]) - s->g[2];
	}
	if (nfree == 2) {
		s->r[2] = -s->theta * (s->z[2] - s->x[2]) - s->g[2];
	}
	if (lq_bmv(s, s->c1h, s->c2h, p1, p2) != 0) {
		return -8;
	}
	LQ_UNROLL
	for (int j = 1; j <= LQ_M; j++) {
		if (j <= col) {
			const double a1 = p1[j];
			const double a2 = s->theta * p2[j];
			if (nfree >= 1) {
				s->r[1] = k1 == 1 ? s->r[1] + s->wy[1][j] * a1 + s->ws[1][j] * a2
				                  : s->r[1] + s->wy[2][j] * a1 + s->ws[2][j] * a2;
			}
			if (nfree == 2) {
				s->r[2] = s->r[2] + s->wy[2][j] * a1 + s->ws[2][j] * a2;
			}
		}
	}
	return 0;
}

/* Triangular solves with the 2col x 2col factor held in wn at block offsets 0 and LQ_M (LINPACK dtrsl over the packed
 * system of the vendored code: unknowns 1..col are the Y half, col+1..2col the S half). */
LQ_FN int lq_dtrsl_wn(const struct lq_state *s, double *b1, double *b2, const int transposed)
{
	const int col = s->col;
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col && s->wn[i][i] == 0.0) return i;
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col && s->wn[LQ_M + i][LQ_M + i] == 0.0) return col + i;
	}
	if (transposed) {
		/* job 11: b[1] /= t11; for j = 2..2col: b[j] = (b[j] - sum_{i<j} t[i][j] b[i]) / t[j][j] */
		b1[1] = lq_div(b1[1], s->wn[1][1]);
		LQ_UNROLL
		for (int j = 2; j <= LQ_M; j++) {
			if (j <= col) {
				double dot = 0.0;
				LQ_UNROLL
				for (int i = 1; i <= LQ_M - 1; i++) {
					if (i <= j - 1) dot += s->wn[i][j] * b1[i];
				}
				b1[j] -= dot;
				b1[j] = lq_div(b1[j], s->wn[j][j]);
			}
		}
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			if (j <= col) {
				const int js = LQ_M + j;
				double dot = 0.0;
				LQ_UNROLL
				for (int i = 1; i <= LQ_M; i++) {
					if (i <= col) dot += s->wn[i][js] * b1[i];
				}
				LQ_UNROLL
				for (int i = 1; i <= LQ_M - 1; i++) {
					if (i <= j - 1) dot += s->wn[LQ_M + i][js] * b2[i];
				}
				b2[j] -= dot;
				b2[j] = lq_div(b2[j], s->wn[js][js]);
			}
		}
	} else {
		/* job 01: b[n] /= t[n][n]; for j = n-1..1: temp = -b[j+1]; b[1..j] += temp * t[1..j][j+1]; b[j] /= t[j][j]
		 * with n = 2col; unknown col+jj lives in b2[jj] */
		LQ_UNROLL
		for (int j = LQ_M; j >= 1; j--) { /* S half: global index col + j */
			if (j == col) {
				b2[j] = lq_div(b2[j], s->wn[LQ_M + j][LQ_M + j]);
			} else if (j < col) {
				const double temp = -b2[j + 1];
				if (temp != 0.0) {
					const int jc = LQ_M + j + 1; /* column of unknown col + j + 1 */
					LQ_UNROLL
					for (int i = 1; i <= LQ_M; i++) {
						if (i <= col) b1[i] += temp * s->wn[i][jc];
					}
					LQ_UNROLL
					for (int i = 1; i <= LQ_M - 1; i++) {
						if (i <= j) b2[i] += temp * s->wn[LQ_M + i][jc];
					}
				}
				b2[j] = lq_div(b2[j], s->wn[LQ_M + j][LQ_M + j]);
			}
		}
		LQ_UNROLL
		for (int j = LQ_M; j >= 1; j--) { /* Y half: global index j; its successor j + 1 is b1[j+1] or, for j = col, b2[1] */
			if (j <= col) {
				const double temp = j == col ? -b2[1] : -b1[j < LQ_M ? j + 1 : j];
				if (temp != 0.0) {
					LQ_UNROLL
					for (int i = 1; i <= LQ_M; i++) {
						if (i <= j) {
							b1[i] += temp * (j == col ? s->wn[i][LQ_M + 1] : s->wn[i][j < LQ_M ? j + 1 : j]);
						}
					}
				}
				b1[j] = lq_div(b1[j], s->wn[j][j]);
			}
		}
	}
	return 0;
}

/* Subspace minimisation with the 2011 projection/backtracking refinement (subalgorithms.c subsm, :1903-2228).
 * On entry z holds the Cauchy point, r the reduced gradient (by free-list position). */
LQ_FN int lq_subsm(struct lq_state *s)
{
	const int col = s->col;
	const int nsub = s->free1 + s->free2;
	const double theta = s->theta;
	double wv1[4], wv2[4];
	double *x = s->z, *d = s->r;
	if (nsub <= 0) {
		return 0;
	}
	const int k1 = s->free1 ? 1 : 2; /* variable at free-list position 1; position 2 (if any) is variable 2 */
	/* wv = W'Z d */
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) {
			double temp1 = 0.0, temp2 = 0.0;
			temp1 += (k1 == 1 ? s->wy[1][i] : s->wy[2][i]) * d[1];
			temp2 += (k1 == 1 ? s->ws[1][i] : s->ws[2][i]) * d[1];
			if (nsub == 2) {
				temp1 += s->wy[2][i] * d[2];
				temp2 += s->ws[2][i] * d[2];
			}
			wv1[i] = temp1;
			wv2[i] = theta * temp2;
		}
	}
	int info = lq_dtrsl_wn(s, wv1, wv2, 1);
	if (info != 0) {
		return info;
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) wv1[i] = -wv1[i];
	}
	info = lq_dtrsl_wn(s, wv1, wv2, 0);
	if (info != 0) {
		return info;
	}
	/* d = (1/theta) d + (1/theta^2) Z'W wv */
	LQ_UNROLL
	for (int jy = 1; jy <= LQ_M; jy++) {
		if (jy <= col) {
			d[1] = d[1] + lq_div((k1 == 1 ? s->wy[1][jy] : s->wy[2][jy]) * wv1[jy], theta) + (k1 == 1 ? s->ws[1][jy] : s->ws[2][jy]) * wv2[jy];
			if (nsub == 2) {
				d[2] = d[2] + lq_div(s->wy[2][jy] * wv1[jy], theta) + s->ws[2][jy] * wv2[jy];
			}
		}
	}
	{
		const double inv = lq_div(1.0, theta);
		d[1] = inv * d[1];
		if (nsub == 2) d[2] = inv * d[2];
	}
	/* projected Newton step */
	int iword = 0;
	const double xp1 = x[1], xp2 = x[2];
	LQ_UNROLL
	for (int i = 1; i <= 2; i++) {
		if (i <= nsub) {
			const int k = i == 1 ? k1 : 2;
			const double dk = d[i];
			if (k == 1) {
				double xk = x[1];
				xk = lq_max(LQ_L1, xk + dk);
				x[1] = lq_min(LQ_U1, xk);
				if (x[1] == LQ_L1 || x[1] == LQ_U1) iword = 1;
			} else {
				double xk = x[2];
				xk = lq_max(LQ_L2, xk + dk);
				x[2] = lq_min(LQ_U2, xk);
				if (x[2] == LQ_L2 || x[2] == LQ_U2) iword = 1;
			}
		}
	}
	if (iword == 0) {
		return 0;
	}
	/* check sign of the directional derivative */
	double dd_p = 0.0;
	dd_p += (x[1] - s->x[1]) * s->g[1];
	dd_p += (x[2] - s->x[2]) * s->g[2];
	if (dd_p > 0.0) {
		x[1] = xp1;
		x[2] = xp2;
		double alpha = 1.0;
		double temp1 = alpha;
		int ibd = 0;
		LQ_UNROLL
		for (int i = 1; i <= 2; i++) {
			if (i <= nsub) {
				const int k = i == 1 ? k1 : 2;
				const double dk = d[i];
				const double xk = k == 1 ? x[1] : x[2];
				const double lk = k == 1 ? LQ_L1 : LQ_L2, uk = k == 1 ? LQ_U1 : LQ_U2;
				if (dk < 0.0) {
					const double temp2 = lk - xk;
					if (temp2 >= 0.0) {
						temp1 = 0.0;
					} else if (dk * alpha < temp2) {
						temp1 = lq_div(temp2, dk);
					}
				} else if (dk > 0.0) {
					const double temp2 = uk - xk;
					if (temp2 <= 0.0) {
						temp1 = 0.0;
					} else if (dk * alpha > temp2) {
						temp1 = lq_div(temp2, dk);
					}
				}
				if (temp1 < alpha) {
					alpha = temp1;
					ibd = i;
				}
			}
		}
		if (alpha < 1.0) {
			const double dk = ibd == 1 ? d[1] : d[2];
			const int k = ibd == 1 ? k1 : 2;
			if (dk > 0.0) {
				if (k == 1) x[1] = LQ_U1; else x[2] = LQ_U2;
				if (ibd == 1) d[1] = 0.0; else d[2] = 0.0;
			} else if (dk < 0.0) {
				if (k == 1) x[1] = LQ_L1; else x[2] = LQ_L2;
				if (ibd == 1) d[1] = 0.0; else d[2] = 0.0;
			}
		}
		LQ_UNROLL
		for (int i = 1; i <= 2; i++) {
			if (i <= nsub) {
				const int k = i == 1 ? k1 : 2;
				if (k == 1) x[1] += alpha * d[i]; else x[2] += alpha * d[i];
			}
		}
	}
	return 0;
}

/* Store the newest correction pair and refresh S'S, S'Y (subalgorithms.c matupd, :1393-1511).  Logical storage:
 * when the memory is full the pairs move down one column and the newest takes column m. */
LQ_FN void lq_matupd(struct lq_state *s, double rr, double dr)
{
	if (s->iupdat <= LQ_M) {
		s->col = s->iupdat;
	} else {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M - 1; j++) {
			s->ws[1][j] = s->ws[1][j + 1]; s->ws[2][j] = s->ws[2][j + 1];
			s->wy[1][j] = s->wy[1][j + 1]; s->wy[2][j] = s->wy[2][j + 1];
		}
	}
	const int col = s->col;
	LQ_UNROLL
	for (int j = 1; j <= LQ_M; j++) {
		if (j == col) {
			s->ws[1][j] = s->d[1]; s->ws[2][j] = s->d[2];
			s->wy[1][j] = s->r[1]; s->wy[2][j] = s->r[2];
		}
	}
	s->theta = lq_div(rr, dr);
	if (s->iupdat > LQ_M) {
		/* move old information */
		LQ_UNROLL
		for (int j = 1; j <= LQ_M - 1; j++) {
			if (j <= col - 1) {
				LQ_UNROLL
				for (int i = 0; i < LQ_M - 1; i++) {
					if (i < j) s->ss[1 + i][j] = s->ss[2 + i][j + 1];
				}
				LQ_UNROLL
				for (int i = 0; i < LQ_M - 1; i++) {
					if (i < col - j) s->sy[j + i][j] = s->sy[j + 1 + i][j + 1];
				}
			}
		}
	}
	LQ_UNROLL
	for (int cc = 1; cc <= LQ_M; cc++) {
		if (cc == col) {
			LQ_UNROLL
			for (int j = 1; j <= LQ_M - 1; j++) {
				if (j <= cc - 1) {
					double a = 0.0, b = 0.0;
					a += s->d[1] * s->wy[1][j];
					a += s->d[2] * s->wy[2][j];
					b += s->ws[1][j] * s->d[1];
					b += s->ws[2][j] * s->d[2];
					s->sy[cc][j] = a;
					s->ss[j][cc] = b;
				}
			}
			if (s->stp == 1.0) {
				s->ss[cc][cc] = s->dtd;
			} else {
				s->ss[cc][cc] = s->stp * s->stp * s->dtd;
			}
			s->sy[cc][cc] = dr;
		}
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		if (i <= col) s->sq[i] = lq_sqrt(s->sy[i][i]);
	}
}

/* T = theta*S'S + L*D^-1*L', Cholesky-factored in place (subalgorithms.c formt, :920-974). */
LQ_FN int lq_formt(struct lq_state *s)
{
	const int col = s->col;
	LQ_UNROLL
	for (int j = 1; j <= LQ_M; j++) {
		if (j <= col) s->wt[1][j] = s->theta * s->ss[1][j];
	}
	LQ_UNROLL
	for (int i = 2; i <= LQ_M; i++) {
		if (i <= col) {
			LQ_UNROLL
			for (int j = 2; j <= LQ_M; j++) {
				if (j >= i && j <= col) {
					const int k1 = (i < j ? i : j) - 1;
					double ddum = 0.0;
					LQ_UNROLL
					for (int k = 1; k <= LQ_M - 1; k++) {
						if (k <= k1) ddum += lq_div(s->sy[i][k] * s->sy[j][k], s->sy[k][k]);
					}
					s->wt[i][j] = ddum + s->theta * s->ss[i][j];
				}
			}
		}
	}
	int info;
	LQ_DPOFA(s->wt, 0, col, info);
	if (info != 0) {
		return -3;
	}
	return 0;
}

/* Safeguarded cubic/quadratic step of Moré & Thuente (linesearch.c dcstep, :485-763).  The four cases of the
 * vendored code build the same cubic-interpolation quantities (theta, s, gamma) from different operands; they are
 * computed once here from case-selected operands — the same operations on the same values in the same order — so
 * that the divisions and the square root exist once in the code instead of four times. */
LQ_FN void lq_dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp,
                      double fp, double dp, int *brackt, double stpmin, double stpmax)
{
	const double sgnd = dp * lq_div(*dx, lq_abs(*dx));
	const int c1 = fp > *fx;                                  /* higher function value: the minimum is bracketed */
	const int c2 = !c1 && sgnd < 0.0;                         /* derivatives of opposite sign: bracketed */
	const int c3 = !c1 && !c2 && lq_abs(dp) < lq_abs(*dx);    /* derivative magnitude decreases */
	const int c4 = !c1 && !c2 && !c3;
	double stpf, stpc, stpq;
	double gamma = 0.0, theta = 0.0, r = 0.0;
	if (!c4 || *brackt) {
		const double num = c4 ? fp - *fy : *fx - fp;
		const double den = c4 ? *sty - *stp : *stp - *stx;
		const double dd = c4 ? *dy : *dx;
		theta = lq_div(num * 3.0, den) + dd + dp;
		const double s = lq_max(lq_max(lq_abs(theta), lq_abs(dd)), lq_abs(dp));
		const double d1 = lq_div(theta, s);
		double rad = d1 * d1 - lq_div(dd, s) * lq_div(dp, s);
		if (c3) {
			rad = lq_max(0.0, rad);
		}
		gamma = s * lq_sqrt(rad);
		const int flip = c1 ? (*stp < *stx) : c4 ? (*stp > *sty) : (*stp > *stx);
		if (flip) {
			gamma = -gamma;
		}
		double p, q;
		if (c1) {
			p = gamma - *dx + theta;
			q = gamma - *dx + gamma + dp;
		} else if (c2) {
			p = gamma - dp + theta;
			q = gamma - dp + gamma + *dx;
		} else if (c3) {
			p = gamma - dp + theta;
			q = gamma + (*dx - dp) + gamma;
		} else {
			p = gamma - dp + theta;
			q = gamma - dp + gamma + *dy;
		}
		r = lq_div(p, q);
	}
	if (c1) {
		stpc = *stx + r * (*stp - *stx);
		stpq = *stx + lq_div(*dx, lq_div(*fx - fp, *stp - *stx) + *dx) / 2.0 * (*stp - *stx);
		if (lq_abs(stpc - *stx) < lq_abs(stpq - *stx)) {
			stpf = stpc;
		} else {
			stpf = stpc + (stpq - stpc) / 2.0;
		}
		*brackt = 1;
	} else if (c2 || c3) {
		stpq = *stp + lq_div(dp, dp - *dx) * (*stx - *stp);
		if (c2) {
			stpc = *stp + r * (*stx - *stp);
			if (lq_abs(stpc - *stp) > lq_abs(stpq - *stp)) {
				stpf = stpc;
			} else {
				stpf = stpq;
			}
			*brackt = 1;
		} else {
			if (r < 0.0 && gamma != 0.0) {
				stpc = *stp + r * (*stx - *stp);
			} else if (*stp > *stx) {
				stpc = stpmax;
			} else {
				stpc = stpmin;
			}
			if (*brackt) {
				if (lq_abs(stpc - *stp) < lq_abs(stpq - *stp)) {
					stpf = stpc;
				} else {
					stpf = stpq;
				}
				if (*stp > *stx) {
					stpf = lq_min(*stp + (*sty - *stp) * 0.66, stpf);
				} else {
					stpf = lq_max(*stp + (*sty - *stp) * 0.66, stpf);
				}
			} else {
				if (lq_abs(stpc - *stp) > lq_abs(stpq - *stp)) {
					stpf = stpc;
				} else {
					stpf = stpq;
				}
				stpf = lq_min(stpmax, stpf);
				stpf = lq_max(stpmin, stpf);
			}
		}
	} else {
		if (*brackt) {
			stpc = *stp + r * (*sty - *stp);
			stpf = stpc;
		} else if (*stp > *stx) {
			stpf = stpmax;
		} else {
			stpf = stpmin;
		}
	}
	if (fp > *fx) {
		*sty = *stp;
		*fy = fp;
		*dy = dp;
	} else {
		if (sgnd < 0.0) {
			*sty = *stx;
			*fy = *fx;
			*dy = *dx;
		}
		*stx = *stp;
		*fx = fp;
		*dx = dp;
	}
	*stp = stpf;
}

/* One reverse-communication turn of the Moré–Thuente search (linesearch.c dcsrch, :161-483). */
LQ_FN void lq_dcsrch_start(struct lq_state *s, double f, double g, const double stp, double stpmax)
{
	s->brackt = 0;
	s->stage = 1;
	s->finit = f;
	s->ginit = g;
	s->gtest = LQ_FTOL * s->ginit;
	s->width = stpmax - LQ_STPMIN;
	s->width1 = s->width / 0.5;
	s->stx = 0.0;
	s->fx = s->finit;
	s->gx = s->ginit;
	s->sty = 0.0;
	s->fy = s->finit;
	s->gy = s->ginit;
	s->stmin = 0.0;
	s->stmax = stp + stp * 4.0;
	s->ls_task = LQ_LS_FG;
}

/* every later turn (task FG) */
LQ_FN void lq_dcsrch(struct lq_state *s, double f, double g, double *stp, double stpmax)
{
	const double ftest = s->finit + *stp * s->gtest;
	if (s->stage == 1 && f <= ftest && g >= 0.0) {
		s->stage = 2;
	}
	int task = LQ_LS_FG;
	if (s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) {
		task = LQ_LS_WARNING;
	}
	if (s->brackt && s->stmax - s->stmin <= LQ_XTOL * s->stmax) {
		task = LQ_LS_WARNING;
	}
	if (*stp == stpmax && f <= ftest && g <= s->gtest) {
		task = LQ_LS_WARNING;
	}
	if (*stp == LQ_STPMIN && (f > ftest || g >= s->gtest)) {
		task = LQ_LS_WARNING;
	}
	if (f <= ftest && lq_abs(g) <= LQ_GTOL * (-s->ginit)) {
		task = LQ_LS_CONVERGENCE;
	}
	if (task != LQ_LS_FG) {
		s->ls_task = task;
		return;
	}
	{
		/* stage 1 with a lower function value that fails the sufficient-decrease test works on the modified
		 * function f(stp) - stp * gtest; one call site serves both forms */
		const int mod = s->stage == 1 && f <= s->fx && f > ftest;
		double fxv = s->fx, fyv = s->fy, gxv = s->gx, gyv = s->gy, fv = f, gv = g;
		if (mod) {
			fv = f - *stp * s->gtest;
			fxv = s->fx - s->stx * s->gtest;
			fyv = s->fy - s->sty * s->gtest;
			gv = g - s->gtest;
			gxv = s->gx - s->gtest;
			gyv = s->gy - s->gtest;
		}
		lq_dcstep(&s->stx, &fxv, &gxv, &s->sty, &fyv, &gyv, stp, fv, gv, &s->brackt, s->stmin, s->stmax);
		if (mod) {
			fxv = fxv + s->stx * s->gtest;
			fyv = fyv + s->sty * s->gtest;
			gxv = gxv + s->gtest;
			gyv = gyv + s->gtest;
		}
		s->fx = fxv; s->fy = fyv; s->gx = gxv; s->gy = gyv;
	}
	if (s->brackt) {
		if (lq_abs(s->sty - s->stx) >= s->width1 * 0.66) {
			*stp = s->stx + (s->sty - s->stx) * 0.5;
		}
		s->width1 = s->width;
		s->width = lq_abs(s->sty - s->stx);
	}
	if (s->brackt) {
		s->stmin = lq_min(s->stx, s->sty);
		s->stmax = lq_max(s->stx, s->sty);
	} else {
		s->stmin = *stp + (*stp - s->stx) * 1.1;
		s->stmax = *stp + (*stp - s->stx) * 4.0;
	}
	*stp = lq_max(*stp, LQ_STPMIN);
	*stp = lq_min(*stp, stpmax);
	if ((s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) ||
	    (s->brackt && s->stmax - s->stmin <= LQ_XTOL * s->stmax)) {
		*stp = s->stx;
	}
	s->ls_task = LQ_LS_FG;
}

/* The reference zeroes its whole workspace before every fit (src/min_saxs.c:217-221); here every element that can be
 * read before it is written is zeroed explicitly (wn1, the pairs), the rest is set where the algorithm sets it. */
LQ_FN void lq_begin(struct lq_state *s, double x1, double x2)
{
	s->x[1] = x1; s->x[2] = x2;
	s->g[1] = 0.0; s->g[2] = 0.0;
	s->f = 0.0;
	LQ_UNROLL
	for (int i = 1; i <= 2; i++) {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) { s->ws[i][j] = 0.0; s->wy[i][j] = 0.0; }
		s->z[i] = s->r[i] = s->d[i] = s->t[i] = 0.0;
	}
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) { s->sy[i][j] = 0.0; s->ss[i][j] = 0.0; s->wt[i][j] = 0.0; }
		s->c1h[i] = 0.0; s->c2h[i] = 0.0;
		s->sq[i] = 0.0;
	}
	LQ_UNROLL
	for (int i = 1; i <= 2 * LQ_M; i++) {
		LQ_UNROLL
		for (int j = 1; j <= 2 * LQ_M; j++) { s->wn[i][j] = 0.0; s->wn1[i][j] = 0.0; }
	}
	s->brackt = 0; s->stage = 0; s->ls_task = LQ_LS_START;
	s->ginit = s->gtest = s->gx = s->gy = s->finit = s->fx = s->fy = 0.0;
	s->stx = s->sty = s->stmin = s->stmax = s->width = s->width1 = 0.0;
	lq_reset_memory(s);
	s->iback = 0;
	s->fold = 0.0; s->gd = 0.0; s->stpmx = 0.0; s->sbgnrm = 0.0;
	s->stp = 0.0; s->gdold = 0.0; s->dtd = 0.0; s->iter = 0; s->nfgv = 0;
	s->free1 = s->free2 = 1;
	s->iw1 = s->iw2 = 0;
	s->ent1 = s->ent2 = s->lv1 = s->lv2 = 0;
	s->ifun = 0; s->wrk = 0;
	s->phase = LQ_PH_INIT;
}

/* One turn of the line search with f, g at the current x (labels 666/556 of mainlb plus the bookkeeping that
 * follows lnsrlb's return, lbfgsb.c:871-915). */
LQ_FN int lq_linesearch_turn(struct lq_state *s, const int first)
{
	int info_ls = 0;
	int search_over = 0;
	{
		double acc = 0.0;
		acc += s->g[1] * s->d[1];
		acc += s->g[2] * s->d[2];
		s->gd = acc;
	}
	if (s->ifun == 0) {
		s->gdold = s->gd;
		if (s->gd >= 0.0) {
			info_ls = -4; /* ascent direction in projection */
		}
	}
	if (info_ls == 0) {
		if (first) {
			lq_dcsrch_start(s, s->f, s->gd, s->stp, s->stpmx); /* the search's START turn: no step is computed */
		} else {
			lq_dcsrch(s, s->f, s->gd, &s->stp, s->stpmx);
		}
		if (s->ls_task == LQ_LS_FG) {
			++s->ifun;
			++s->nfgv;
			s->iback = s->ifun - 1;
			if (s->stp == 1.0) {
				s->x[1] = s->z[1];
				s->x[2] = s->z[2];
			} else {
				s->x[1] = s->stp * s->d[1] + s->t[1];
				s->x[2] = s->stp * s->d[2] + s->t[2];
			}
		} else {
			search_over = 1;
		}
	}
	if (info_ls != 0 || s->iback >= 20) {
		/* restore the previous iterate */
		s->x[1] = s->t[1]; s->x[2] = s->t[2];
		s->g[1] = s->r[1]; s->g[2] = s->r[2];
		s->f = s->fold;
		if (s->col == 0) {
			if (info_ls == 0) {
				--s->nfgv;
				--s->ifun;
				--s->iback;
			}
			++s->iter;
			return LQ_PH_DONE;
		}
		if (info_ls == 0) {
			--s->nfgv;
		}
		lq_reset_memory(s);
		return LQ_PH_B_NEW_ITER;
	}
	return search_over ? LQ_PH_B_ACCEPTED : LQ_PH_LINESEARCH;
}

/* Part A — right after an objective evaluation: the line-search turn. */
LQ_FN int lq_step_a(struct lq_state *s)
{
	if (s->phase == LQ_PH_LINESEARCH) {
		s->phase = lq_linesearch_turn(s, 0);
		if (s->phase == LQ_PH_LINESEARCH) {
			return LQ_NEED_EVAL;
		}
		return s->phase == LQ_PH_DONE ? LQ_DONE : LQ_NEED_B;
	}
	if (s->phase == LQ_PH_FIRST_EVAL) {
		s->phase = LQ_PH_B_FIRST;
		return LQ_NEED_B;
	}
	if (s->phase == LQ_PH_INIT) {
		/* active(): project the start into the box (subalgorithms.c:7-118) */
		LQ_UNROLL
		for (int i = 1; i <= 2; i++) {
			if (s->x[i] <= LQ_LO(i)) {
				s->x[i] = LQ_LO(i);
			} else if (s->x[i] >= LQ_UP(i)) {
				s->x[i] = LQ_UP(i);
			}
		}
		s->iw1 = (LQ_U1 - LQ_L1 <= 0.0) ? 3 : 0;
		s->iw2 = (LQ_U2 - LQ_L2 <= 0.0) ? 3 : 0;
		s->phase = LQ_PH_FIRST_EVAL;
		return LQ_NEED_EVAL;
	}
	return s->phase == LQ_PH_DONE ? LQ_DONE : LQ_NEED_B;
}

/* Part B — the iteration boundary (lbfgsb.c mainlb, labels 222..777): convergence tests, BFGS update, generalised
 * Cauchy point, subspace minimisation, line-search set-up and its first turn.  Returns LQ_NEED_EVAL or LQ_DONE. */
LQ_FN int lq_step_b_body(struct lq_state *s, const double pgtol, const double tol);
LQ_FN int lq_step_b(struct lq_state *s, const double pgtol, const double tol)
{
	const int r = lq_step_b_body(s, pgtol, tol);
	LQ_UNROLL
	for (int i = 1; i <= LQ_M; i++) {
		LQ_UNROLL
		for (int j = 1; j <= LQ_M; j++) {
			s->wt[i][j] = 0.0; /* dead until the next boundary forms it again */
		}
	}
	return r;
}

LQ_FN int lq_step_b_body(struct lq_state *s, const double pgtol, const double tol)
{
	/* entry: 0 = test convergence of a first evaluation, 1 = accepted iterate, 2 = new iteration */
	int entry = s->phase == LQ_PH_B_NEW_ITER ? 2 : s->phase == LQ_PH_B_ACCEPTED ? 1 : s->phase == LQ_PH_B_FIRST ? 0 : -1;
	if (entry < 0) {
		return s->phase == LQ_PH_DONE ? LQ_DONE : LQ_NEED_EVAL;
	}
	/* The vendored code is a web of gotos (222 = new iteration, 777 = accepted); a memory reset after a failed
	 * factorisation restarts the iteration with col = 0, which cannot fail again, so two passes bound the loop;
	 * the pass counter is a safety net only. */
	for (int pass = 0; pass < 64; pass++) {
		if (entry == 0) {
			s->nfgv = 1;
			s->sbgnrm = lq_projgr(s);
			if (s->sbgnrm <= pgtol) {
				s->phase = LQ_PH_DONE;
				return LQ_DONE;
			}
			entry = 2;
		}
		if (entry == 1) {
			++s->iter;
			s->sbgnrm = lq_projgr(s);
			if (s->sbgnrm <= pgtol) {
				s->phase = LQ_PH_DONE;
				return LQ_DONE;
			}
			{
				const double ddum = lq_max(lq_max(lq_abs(s->fold), lq_abs(s->f)), 1.0);
				if (s->fold - s->f <= tol * ddum) {
					s->phase = LQ_PH_DONE;
					return LQ_DONE;
				}
			}
			/* ---- BFGS update ---- */
			s->r[1] = s->g[1] - s->r[1];
			s->r[2] = s->g[2] - s->r[2];
			double rr = 0.0;
			rr += s->r[1] * s->r[1];
			rr += s->r[2] * s->r[2];
			double dr, ddum2;
			if (s->stp == 1.0) {
				dr = s->gd - s->gdold;
				ddum2 = -s->gdold;
			} else {
				dr = (s->gd - s->gdold) * s->stp;
				s->d[1] = s->stp * s->d[1];
				s->d[2] = s->stp * s->d[2];
				ddum2 = -s->gdold * s->stp;
			}
			if (dr <= LQ_EPSMCH * ddum2) {
				s->updatd = 0; /* skip the update */
			} else {
				s->updatd = 1;
				++s->iupdat;
				lq_matupd(s, rr, dr);
			}
			/* The vendored code forms T only after an update and keeps its factor otherwise.  Here it is formed at every
			 * boundary that has corrections: without an update theta, S'S, S'Y and col are unchanged, so the same
			 * factor comes out — and wt need not survive from one iteration to the next (lq_step_b clears it on the
			 * way out: 9 doubles less to keep per fit across the objective). */
			if (s->col > 0 && lq_formt(s) != 0) {
				lq_reset_memory(s);
			}
			entry = 2;
		}
		/* ---- new iteration (label 222) ---- */
		if (lq_cauchy(s) != 0) {
			lq_reset_memory(s);
			continue;
		}
		lq_freev(s);
		const int nfree = s->free1 + s->free2;
		if (nfree != 0 && s->col != 0) {
			int info = lq_formk(s);
			if (info != 0) {
				lq_reset_memory(s);
				continue;
			}
			info = lq_cmprlb(s);
			if (info == 0) {
				info = lq_subsm(s);
			}
			if (info != 0) {
				lq_reset_memory(s);
				continue;
			}
		}
		/* ---- line search along d = z - x (linesearch.c lnsrlb, :5-159) ---- */
		s->d[1] = s->z[1] - s->x[1];
		s->d[2] = s->z[2] - s->x[2];
		{
			double acc = 0.0;
			acc += s->d[1] * s->d[1];
			acc += s->d[2] * s->d[2];
			s->dtd = acc;
		}
		s->stpmx = 1e10;
		if (s->iter == 0) {
			s->stpmx = 1.0;
		} else {
			LQ_UNROLL
			for (int i = 1; i <= 2; i++) {
				const double a1 = s->d[i];
				if (a1 < 0.0) {
					const double a2 = LQ_LO(i) - s->x[i];
					if (a2 >= 0.0) {
						s->stpmx = 0.0;
					} else if (a1 * s->stpmx < a2) {
						s->stpmx = lq_div(a2, a1);
					}
				} else if (a1 > 0.0) {
					const double a2 = LQ_UP(i) - s->x[i];
					if (a2 <= 0.0) {
						s->stpmx = 0.0;
					} else if (a1 * s->stpmx > a2) {
						s->stpmx = lq_div(a2, a1);
					}
				}
			}
		}
		s->stp = 1.0;
		s->t[1] = s->x[1]; s->t[2] = s->x[2];
		s->r[1] = s->g[1]; s->r[2] = s->g[2];
		s->fold = s->f;
		s->ifun = 0;
		s->iback = 0;
		s->ls_task = LQ_LS_START;
		/* first turn of the search: uses f, g at the current iterate, no new evaluation needed */
		s->phase = lq_linesearch_turn(s, 1);
		if (s->phase == LQ_PH_LINESEARCH) {
			return LQ_NEED_EVAL;
		}
		if (s->phase == LQ_PH_B_NEW_ITER) {
			entry = 2;
			continue;
		}
		if (s->phase == LQ_PH_DONE) {
			return LQ_DONE;
		}
		entry = 1; /* LQ_PH_B_ACCEPTED cannot follow a START turn, kept for completeness */
	}
	s->phase = LQ_PH_DONE;
	return LQ_DONE;
}

/* Serial composition of the two parts (host harness, single fits). */
LQ_FN int lq_step(struct lq_state *s, const double pgtol, const double tol)
{
	int r = lq_step_a(s);
	if (r == LQ_NEED_B) {
		r = lq_step_b(s, pgtol, tol);
	}
	return r;
}

#endif /* SXS_LBFGSB_LEAN_H */
