/* Per-thread bound-constrained limited-memory BFGS for the (c1, c2) fit.
 *
 * Algorithm: L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu 1995; Morales & Nocedal 2011) exactly as the
 * reference drives it from src/min_saxs.c:196-259 — n = 2 variables, m = 3 corrections, both
 * variables boxed, factr = 1e7, pgtol = 1e-5, start (1.0, 0.0).  The reference vendors the f2c'd
 * "L-BFGS-B-C" (lbfgsb/src/lbfgsb.c, subalgorithms.c, linesearch.c, linpack.c, miniCBLAS.c) whose
 * state lives in function-local statics and a 371 KB workspace; here the whole optimiser state is
 * one small struct in thread-local memory so that every GPU thread runs its own fit.
 *
 * Parity notes: the minimiser stops mid-convergence, so the iterates must follow
 * the reference's floating-point path, not just its mathematics.  Every accumulation below keeps
 * the reference's evaluation order (its BLAS-1 loops are sequential sums), and the translation
 * unit that includes this file is compiled with -fmad=false so nothing is contracted into FMAs.
 *
 * The header is plain C/C++ and carries no CUDA dependency besides the SXS_HD qualifier, so the
 * CPU test-suite can compile it with gcc and compare iterates against the reference on the host.
 *
 * Third-party notice.  The optimiser below restates, function by function and in the same floating-point
 * order, L-BFGS-B 3.0 (C. Zhu, R. Byrd, J. Nocedal, J. L. Morales; released under the "New BSD License")
 * in the C version "L-BFGS-B-C" that libfmftsaxs vendors under lbfgsb/ (lbfgsb/LICENSE), whose notice is
 * carried here as its terms require:
 *
 *   BSD 3-clause license
 *   Copyright (c) 2015, Stephen Becker
 *   All rights reserved.
 *
 *   Redistribution and use in source and binary forms, with or without
 *   modification, are permitted provided that the following conditions are met:
 *
 *   * Redistributions of source code must retain the above copyright notice, this
 *     list of conditions and the following disclaimer.
 *
 *   * Redistributions in binary form must reproduce the above copyright notice,
 *     this list of conditions and the following disclaimer in the documentation
 *     and/or other materials provided with the distribution.
 *
 *   * Neither the name of the copyright holder nor the names of its
 *     contributors may be used to endorse or promote products derived from
 *     this software without specific prior written permission.
 *
 *   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS"
 *   AND ANY EXPRESS OR IMPLIED WARRANTIES, INCLUDING, BUT NOT LIMITED TO, THE
 *   IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE
 *   DISCLAIMED. IN NO EVENT SHALL THE COPYRIGHT HOLDER OR CONTRIBUTORS BE LIABLE
 *   FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL
 *   DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR
 *   SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER
 *   CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER IN CONTRACT, STRICT LIABILITY,
 *   OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE USE
 *   OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
 */
#ifndef SXS_LBFGSB_N2M3_H
#define SXS_LBFGSB_N2M3_H

#include <math.h>

#ifndef SXS_HD
#ifdef __CUDACC__
#define SXS_HD __host__ __device__ __forceinline__
#else
#define SXS_HD static inline
#endif
#endif

/* The optimiser is control-flow heavy and tiny in arithmetic.  On the GPU it is kept as real
 * (non-inlined) functions with rolled loops: fully inlined and unrolled it compiles to ~20 000 SASS
 * instructions (330 KB), and the divergent lanes of a warp then stall on instruction-cache misses
 * (ncu: 63 % of warp cycles in "no instruction"); the objective evaluation stays inlined. */
#ifdef __CUDACC__
#ifndef LB_FN
#define LB_FN __host__ __device__ __noinline__
#endif
#ifndef LB_NOUNROLL
#define LB_NOUNROLL _Pragma("unroll 1")
#endif
#else
#define LB_FN static
#define LB_NOUNROLL
#endif

#define LB_N 2
#define LB_M 3
#define LB_M2 (2 * LB_M)

/* lbfgsb/src/lbfgsb.h:213-222 */
#define LB_FTOL 1.0e-3
#define LB_GTOL 0.9
#define LB_XTOL 0.1
#define LB_STPMIN 0.0
/* float.h DBL_EPSILON, lbfgsb/src/lbfgsb.c mainlb START branch */
#define LB_EPSMCH 2.220446049250313e-16

enum lb_search_task { LS_START = 0, LS_FG, LS_CONVERGENCE, LS_WARNING };

/* 1-based storage (index 0 unused) mirrors the Fortran-style indexing of the algorithm papers. */
struct lb_state {
	double x[LB_N + 1], l[LB_N + 1], u[LB_N + 1], g[LB_N + 1];
	double f;
	double ws[LB_N + 1][LB_M + 1]; /* S: column j is the j-th stored step */
	double wy[LB_N + 1][LB_M + 1]; /* Y */
	double sy[LB_M + 1][LB_M + 1];
	double ss[LB_M + 1][LB_M + 1];
	double wt[LB_M + 1][LB_M + 1];
	double wn[LB_M2 + 1][LB_M2 + 1];
	double wn1[LB_M2 + 1][LB_M2 + 1];
	double z[LB_N + 1], r[LB_N + 1], d[LB_N + 1], t[LB_N + 1], xp[LB_N + 1];
	double wa[8 * LB_M + 1];
	int index[LB_N + 1], iwhere[LB_N + 1], indx2[LB_N + 1];
	double theta, fold, dnorm, gd, gdold, stp, stpmx, sbgnrm, dtd, tol;
	int col, head, itail, iupdat, updatd, iback, ifun, iter, nfgv, nfree, nact, ileave, nenter, nseg, wrk;
	int phase; /* enum lb_phase: where lb_step() resumes */
	/* Moré–Thuente search state */
	int brackt, stage, ls_task;
	double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

SXS_HD double lb_abs(double a) { return a >= 0 ? a : -a; }
SXS_HD double lb_max(double a, double b) { return a >= b ? a : b; }
SXS_HD double lb_min(double a, double b) { return a <= b ? a : b; }

SXS_HD void lb_reset_memory(struct lb_state *s)
{
	s->col = 0;
	s->head = 1;
	s->theta = 1.0;
	s->iupdat = 0;
	s->updatd = 0;
}

/* Cholesky factor (upper) of the n x n block of `a` whose (1,1) sits at row/column off+1;
 * LINPACK dpofa, lbfgsb/src/linpack.c:11-109.  a is a row-major [ld][ld] array addressed a[row*ld + col]
 * (1-based rows and columns, like the state struct).  One instance serves the m x m and 2m x 2m
 * matrices: code size matters more than the constant stride here (see the note above LB_FN). */
LB_FN int lb_dpofa(double *a, int ld, int off, int n)
{
	LB_NOUNROLL
	for (int j = 1; j <= n; j++) {
		double sacc = 0.0;
		LB_NOUNROLL
		for (int k = 1; k <= j - 1; k++) {
			double dot = 0.0;
			LB_NOUNROLL
			for (int i = 1; i <= k - 1; i++) {
				dot += a[(off + i) * ld + off + k] * a[(off + i) * ld + off + j];
			}
			double tt = a[(off + k) * ld + off + j] - dot;
			tt /= a[(off + k) * ld + off + k];
			a[(off + k) * ld + off + j] = tt;
			sacc += tt * tt;
		}
		sacc = a[(off + j) * ld + off + j] - sacc;
		if (sacc <= 0.0) {
			return j;
		}
		a[(off + j) * ld + off + j] = sqrt(sacc);
	}
	return 0;
}

/* Triangular solves with an upper-triangular factor stored in t[row*ld + col] (LINPACK dtrsl,
 * lbfgsb/src/linpack.c:111-297): job 11 solves trans(T) x = b, job 01 solves T x = b. */
LB_FN int lb_dtrsl(const double *t, int ld, int n, double *b, int job)
{
	LB_NOUNROLL
	for (int i = 1; i <= n; i++) {
		if (t[i * ld + i] == 0.0) {
			return i;
		}
	}
	if (job == 11) {
		b[1] /= t[1 * ld + 1];
		LB_NOUNROLL
		for (int j = 2; j <= n; j++) {
			double dot = 0.0;
			LB_NOUNROLL
			for (int i = 1; i <= j - 1; i++) {
				dot += t[i * ld + j] * b[i];
			}
			b[j] -= dot;
			b[j] /= t[j * ld + j];
		}
	} else { /* job == 01 */
		b[n] /= t[n * ld + n];
		LB_NOUNROLL
		for (int jj = 2; jj <= n; jj++) {
			int j = n - jj + 1;
			double temp = -b[j + 1];
			if (temp != 0.0) {
				LB_NOUNROLL
				for (int i = 1; i <= j; i++) {
					b[i] += temp * t[i * ld + j + 1];
				}
			}
			b[j] /= t[j * ld + j];
		}
	}
	return 0;
}

/* Product of the 2m x 2m middle matrix of the compact L-BFGS formula with a 2*col vector
 * (subalgorithms.c bmv, :120-259). */
LB_FN int lb_bmv(const struct lb_state *s, const double *v, double *p)
{
	const int col = s->col;
	if (col == 0) {
		return 0;
	}
	/* solve [ D^(1/2)  O ] [ p1 ] = [ v1 ]
	 *       [ -L*D^(-1/2) J ] [ p2 ]   [ v2 ]  */
	p[col + 1] = v[col + 1];
	LB_NOUNROLL
	for (int i = 2; i <= col; i++) {
		int i2 = col + i;
		double sum = 0.0;
		LB_NOUNROLL
		for (int k = 1; k <= i - 1; k++) {
			sum += s->sy[i][k] * v[k] / s->sy[k][k];
		}
		p[i2] = v[i2] + sum;
	}
	int info = lb_dtrsl(&s->wt[0][0], LB_M + 1, col, &p[col], 11);
	if (info != 0) {
		return info;
	}
	LB_NOUNROLL
	for (int i = 1; i <= col; i++) {
		p[i] = v[i] / sqrt(s->sy[i][i]);
	}
	/* solve [ -D^(1/2)  D^(-1/2)*L' ] [ p1 ] = [ p1 ]
	 *       [ 0         J'          ] [ p2 ]   [ p2 ]  */
	info = lb_dtrsl(&s->wt[0][0], LB_M + 1, col, &p[col], 1);
	if (info != 0) {
		return info;
	}
	LB_NOUNROLL
	for (int i = 1; i <= col; i++) {
		p[i] = -p[i] / sqrt(s->sy[i][i]);
	}
	LB_NOUNROLL
	for (int i = 1; i <= col; i++) {
		double sum = 0.0;
		LB_NOUNROLL
		for (int k = i + 1; k <= col; k++) {
			sum += s->sy[k][i] * p[col + k] / s->sy[i][i];
		}
		p[i] += sum;
	}
	return 0;
}

/* Heap extraction of the next breakpoint (subalgorithms.c hpsolb, :1816-1901). */
LB_FN void lb_hpsolb(int n, double *t, int *iorder, int iheap)
{
	if (iheap == 0) {
		LB_NOUNROLL
		for (int k = 2; k <= n; k++) {
			double ddum = t[k];
			int indxin = iorder[k];
			int i = k;
			while (i > 1) {
				int j = i / 2;
				if (ddum < t[j]) {
					t[i] = t[j];
					iorder[i] = iorder[j];
					i = j;
				} else {
					break;
				}
			}
			t[i] = ddum;
			iorder[i] = indxin;
		}
	}
	if (n > 1) {
		int i = 1;
		double out = t[1];
		int indxou = iorder[1];
		double ddum = t[n];
		int indxin = iorder[n];
		LB_NOUNROLL
		for (;;) {
			int j = i + i;
			if (j <= n - 1) {
				if (t[j + 1] < t[j]) {
					j++;
				}
				if (t[j] < ddum) {
					t[i] = t[j];
					iorder[i] = iorder[j];
					i = j;
					continue;
				}
			}
			break;
		}
		t[i] = ddum;
		iorder[i] = indxin;
		t[n] = out;
		iorder[n] = indxou;
	}
}

/* Projected-gradient sup-norm (subalgorithms.c projgr, :1513-1560); both variables are boxed. */
LB_FN double lb_projgr(const struct lb_state *s)
{
	double sbgnrm = 0.0;
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		double gi = s->g[i];
		if (gi < 0.0) {
			gi = lb_max(s->x[i] - s->u[i], gi);
		} else {
			gi = lb_min(s->x[i] - s->l[i], gi);
		}
		sbgnrm = lb_max(sbgnrm, lb_abs(gi));
	}
	return sbgnrm;
}

/* Generalised Cauchy point (subalgorithms.c cauchy, :261-818).  Workspace split of wa as in
 * mainlb: p = wa[1..2m], c = wa[2m+1..4m], wbp = wa[4m+1..6m], v = wa[6m+1..8m].
 * Breakpoint times share storage with s->t, the order array with s->indx2 (as mainlb passes them). */
LB_FN int lb_cauchy(struct lb_state *s)
{
	double *p = &s->wa[0], *c = &s->wa[2 * LB_M], *wbp = &s->wa[4 * LB_M], *v = &s->wa[6 * LB_M];
	double *t = s->t, *d = s->d, *xcp = s->z;
	int *iorder = s->indx2, *iwhere = s->iwhere;
	const int col = s->col;
	const double theta = s->theta;

	if (s->sbgnrm <= 0.0) {
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			xcp[i] = s->x[i];
		}
		return 0;
	}
	int bnded = 1;
	int nfree = LB_N + 1;
	int nbreak = 0;
	int ibkmin = 0;
	double bkmin = 0.0;
	const int col2 = 2 * col;
	double f1 = 0.0;
	double tl = 0.0, tu = 0.0;

	LB_NOUNROLL
	for (int i = 1; i <= col2; i++) {
		p[i] = 0.0;
	}
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		double neggi = -s->g[i];
		if (iwhere[i] != 3 && iwhere[i] != -1) {
			tl = s->x[i] - s->l[i];
			tu = s->u[i] - s->x[i];
			int xlower = tl <= 0.0;
			int xupper = tu <= 0.0;
			iwhere[i] = 0;
			if (xlower) {
				if (neggi <= 0.0) {
					iwhere[i] = 1;
				}
			} else if (xupper) {
				if (neggi >= 0.0) {
					iwhere[i] = 2;
				}
			} else {
				if (lb_abs(neggi) <= 0.0) {
					iwhere[i] = -3;
				}
			}
		}
		int pointr = s->head;
		if (iwhere[i] != 0 && iwhere[i] != -1) {
			d[i] = 0.0;
		} else {
			d[i] = neggi;
			f1 -= neggi * neggi;
			LB_NOUNROLL
			for (int j = 1; j <= col; j++) {
				p[j] += s->wy[i][pointr] * neggi;
				p[col + j] += s->ws[i][pointr] * neggi;
				pointr = pointr % LB_M + 1;
			}
			if (neggi < 0.0) {
				++nbreak;
				iorder[nbreak] = i;
				t[nbreak] = tl / (-neggi);
				if (nbreak == 1 || t[nbreak] < bkmin) {
					bkmin = t[nbreak];
					ibkmin = nbreak;
				}
			} else if (neggi > 0.0) {
				++nbreak;
				iorder[nbreak] = i;
				t[nbreak] = tu / neggi;
				if (nbreak == 1 || t[nbreak] < bkmin) {
					bkmin = t[nbreak];
					ibkmin = nbreak;
				}
			} else {
				--nfree;
				iorder[nfree] = i;
				if (lb_abs(neggi) > 0.0) {
					bnded = 0;
				}
			}
		}
	}
	if (theta != 1.0) {
		LB_NOUNROLL
		for (int j = 1; j <= col; j++) {
			p[col + j] = theta * p[col + j];
		}
	}
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		xcp[i] = s->x[i];
	}
	if (nbreak == 0 && nfree == LB_N + 1) {
		return 0;
	}
	LB_NOUNROLL
	for (int j = 1; j <= col2; j++) {
		c[j] = 0.0;
	}
	double f2 = -theta * f1;
	const double f2_org = f2;
	if (col > 0) {
		int info = lb_bmv(s, p, v);
		if (info != 0) {
			return info;
		}
		double dot = 0.0;
		LB_NOUNROLL
		for (int j = 1; j <= col2; j++) {
			dot += v[j] * p[j];
		}
		f2 -= dot;
	}
	double dtm = -f1 / f2;
	double tsum = 0.0;
	s->nseg = 1;
	int skip_to_end = 0; /* the reference's label 999: all variables hit their bounds */

	if (nbreak != 0) {
		int nleft = nbreak;
		int iter = 1;
		double tj = 0.0;
		LB_NOUNROLL
		for (;;) {
			double tj0 = tj;
			int ibp;
			if (iter == 1) {
				tj = bkmin;
				ibp = iorder[ibkmin];
			} else {
				if (iter == 2) {
					if (ibkmin != nbreak) {
						t[ibkmin] = t[nbreak];
						iorder[ibkmin] = iorder[nbreak];
					}
				}
				lb_hpsolb(nleft, t, iorder, iter - 2);
				tj = t[nleft];
				ibp = iorder[nleft];
			}
			double dt = tj - tj0;
			if (dtm < dt) {
				break;
			}
			tsum += dt;
			--nleft;
			++iter;
			double dibp = d[ibp];
			d[ibp] = 0.0;
			double zibp;
			if (dibp > 0.0) {
				zibp = s->u[ibp] - s->x[ibp];
				xcp[ibp] = s->u[ibp];
				iwhere[ibp] = 2;
			} else {
				zibp = s->l[ibp] - s->x[ibp];
				xcp[ibp] = s->l[ibp];
				iwhere[ibp] = 1;
			}
			if (nleft == 0 && nbreak == LB_N) {
				dtm = dt;
				skip_to_end = 1;
				break;
			}
			++s->nseg;
			double dibp2 = dibp * dibp;
			f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
			f2 -= theta * dibp2;
			if (col > 0) {
				if (dt != 0.0) {
					LB_NOUNROLL
					for (int j = 1; j <= col2; j++) {
						c[j] += dt * p[j];
					}
				}
				int pointr = s->head;
				LB_NOUNROLL
				for (int j = 1; j <= col; j++) {
					wbp[j] = s->wy[ibp][pointr];
					wbp[col + j] = theta * s->ws[ibp][pointr];
					pointr = pointr % LB_M + 1;
				}
				int info = lb_bmv(s, wbp, v);
				if (info != 0) {
					return info;
				}
				double wmc = 0.0, wmp = 0.0, wmw = 0.0;
				LB_NOUNROLL
				for (int j = 1; j <= col2; j++) {
					wmc += c[j] * v[j];
				}
				LB_NOUNROLL
				for (int j = 1; j <= col2; j++) {
					wmp += p[j] * v[j];
				}
				LB_NOUNROLL
				for (int j = 1; j <= col2; j++) {
					wmw += wbp[j] * v[j];
				}
				double mdibp = -dibp;
				if (mdibp != 0.0) {
					LB_NOUNROLL
					for (int j = 1; j <= col2; j++) {
						p[j] += mdibp * wbp[j];
					}
				}
				f1 += dibp * wmc;
				f2 = f2 + dibp * 2.0 * wmp - dibp2 * wmw;
			}
			f2 = lb_max(LB_EPSMCH * f2_org, f2);
			if (nleft > 0) {
				dtm = -f1 / f2;
				continue;
			} else if (bnded) {
				f1 = 0.0;
				f2 = 0.0;
				dtm = 0.0;
			} else {
				dtm = -f1 / f2;
			}
			break;
		}
	}
	if (!skip_to_end) {
		if (dtm <= 0.0) {
			dtm = 0.0;
		}
		tsum += dtm;
		if (tsum != 0.0) {
			LB_NOUNROLL
			for (int i = 1; i <= LB_N; i++) {
				xcp[i] += tsum * d[i];
			}
		}
	}
	if (col > 0 && dtm != 0.0) {
		LB_NOUNROLL
		for (int j = 1; j <= col2; j++) {
			c[j] += dtm * p[j];
		}
	}
	return 0;
}

/* Free/active bookkeeping at the Cauchy point (subalgorithms.c freev, :1741-1814). */
LB_FN void lb_freev(struct lb_state *s)
{
	int *index = s->index, *indx2 = s->indx2, *iwhere = s->iwhere;
	s->nenter = 0;
	s->ileave = LB_N + 1;
	if (s->iter > 0) {
		LB_NOUNROLL
		for (int i = 1; i <= s->nfree; i++) {
			int k = index[i];
			if (iwhere[k] > 0) {
				--s->ileave;
				indx2[s->ileave] = k;
			}
		}
		LB_NOUNROLL
		for (int i = s->nfree + 1; i <= LB_N; i++) {
			int k = index[i];
			if (iwhere[k] <= 0) {
				++s->nenter;
				indx2[s->nenter] = k;
			}
		}
	}
	s->wrk = (s->ileave < LB_N + 1) || (s->nenter > 0) || s->updatd;
	s->nfree = 0;
	int iact = LB_N + 1;
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		if (iwhere[i] <= 0) {
			++s->nfree;
			index[s->nfree] = i;
		} else {
			--iact;
			index[iact] = i;
		}
	}
}

/* LEL^T factorisation of the indefinite subspace matrix (subalgorithms.c formk, :820-1303),
 * including its incremental update of the lower triangle of N kept in wn1. */
LB_FN int lb_formk(struct lb_state *s)
{
	const int col = s->col, head = s->head, nsub = s->nfree;
	const int *ind = s->index, *indx2 = s->indx2;
	double (*wn1)[LB_M2 + 1] = s->wn1;
	double (*wn)[LB_M2 + 1] = s->wn;
	int upcl;

	if (s->updatd) {
		if (s->iupdat > LB_M) {
			/* shift old part of WN1 */
			LB_NOUNROLL
			for (int jy = 1; jy <= LB_M - 1; jy++) {
				int js = LB_M + jy;
				LB_NOUNROLL
				for (int i = 0; i < LB_M - jy; i++) {
					wn1[jy + i][jy] = wn1[jy + 1 + i][jy + 1];
				}
				LB_NOUNROLL
				for (int i = 0; i < LB_M - jy; i++) {
					wn1[js + i][js] = wn1[js + 1 + i][js + 1];
				}
				LB_NOUNROLL
				for (int i = 0; i < LB_M - 1; i++) {
					wn1[LB_M + 1 + i][jy] = wn1[LB_M + 2 + i][jy + 1];
				}
			}
		}
		/* put new rows in blocks (1,1), (2,1) and (2,2) */
		const int pbegin = 1, pend = nsub, dbegin = nsub + 1, dend = LB_N;
		int iy = col;
		int is = LB_M + col;
		int ipntr = head + col - 1;
		if (ipntr > LB_M) {
			ipntr -= LB_M;
		}
		int jpntr = head;
		LB_NOUNROLL
		for (int jy = 1; jy <= col; jy++) {
			int js = LB_M + jy;
			double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0;
			LB_NOUNROLL
			for (int k = pbegin; k <= pend; k++) {
				int k1 = ind[k];
				temp1 += s->wy[k1][ipntr] * s->wy[k1][jpntr];
			}
			LB_NOUNROLL
			for (int k = dbegin; k <= dend; k++) {
				int k1 = ind[k];
				temp2 += s->ws[k1][ipntr] * s->ws[k1][jpntr];
				temp3 += s->ws[k1][ipntr] * s->wy[k1][jpntr];
			}
			wn1[iy][jy] = temp1;
			wn1[is][js] = temp2;
			wn1[is][jy] = temp3;
			jpntr = jpntr % LB_M + 1;
		}
		/* put new column in block (2,1) */
		int jy = col;
		jpntr = head + col - 1;
		if (jpntr > LB_M) {
			jpntr -= LB_M;
		}
		ipntr = head;
		LB_NOUNROLL
		for (int i = 1; i <= col; i++) {
			int is2 = LB_M + i;
			double temp3 = 0.0;
			LB_NOUNROLL
			for (int k = pbegin; k <= pend; k++) {
				int k1 = ind[k];
				temp3 += s->ws[k1][ipntr] * s->wy[k1][jpntr];
			}
			ipntr = ipntr % LB_M + 1;
			wn1[is2][jy] = temp3;
		}
		upcl = col - 1;
	} else {
		upcl = col;
	}
	/* modify the old parts in blocks (1,1) and (2,2) due to changes in the set of free variables */
	int ipntr = head;
	LB_NOUNROLL
	for (int iy = 1; iy <= upcl; iy++) {
		int is = LB_M + iy;
		int jpntr = head;
		LB_NOUNROLL
		for (int jy = 1; jy <= iy; jy++) {
			int js = LB_M + jy;
			double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0, temp4 = 0.0;
			LB_NOUNROLL
			for (int k = 1; k <= s->nenter; k++) {
				int k1 = indx2[k];
				temp1 += s->wy[k1][ipntr] * s->wy[k1][jpntr];
				temp2 += s->ws[k1][ipntr] * s->ws[k1][jpntr];
			}
			LB_NOUNROLL
			for (int k = s->ileave; k <= LB_N; k++) {
				int k1 = indx2[k];
				temp3 += s->wy[k1][ipntr] * s->wy[k1][jpntr];
				temp4 += s->ws[k1][ipntr] * s->ws[k1][jpntr];
			}
			wn1[iy][jy] = wn1[iy][jy] + temp1 - temp3;
			wn1[is][js] = wn1[is][js] - temp2 + temp4;
			jpntr = jpntr % LB_M + 1;
		}
		ipntr = ipntr % LB_M + 1;
	}
	/* modify the old parts in block (2,1) */
	ipntr = head;
	LB_NOUNROLL
	for (int is = LB_M + 1; is <= LB_M + upcl; is++) {
		int jpntr = head;
		LB_NOUNROLL
		for (int jy = 1; jy <= upcl; jy++) {
			double temp1 = 0.0, temp3 = 0.0;
			LB_NOUNROLL
			for (int k = 1; k <= s->nenter; k++) {
				int k1 = indx2[k];
				temp1 += s->ws[k1][ipntr] * s->wy[k1][jpntr];
			}
			LB_NOUNROLL
			for (int k = s->ileave; k <= LB_N; k++) {
				int k1 = indx2[k];
				temp3 += s->ws[k1][ipntr] * s->wy[k1][jpntr];
			}
			if (is <= jy + LB_M) {
				wn1[is][jy] = wn1[is][jy] + temp1 - temp3;
			} else {
				wn1[is][jy] = wn1[is][jy] - temp1 + temp3;
			}
			jpntr = jpntr % LB_M + 1;
		}
		ipntr = ipntr % LB_M + 1;
	}
	/* form the upper triangle of WN = [D+Y'ZZ'Y/theta   -L_a'+R_z' ; -L_a+R_z   S'AA'S*theta] */
	const double theta = s->theta;
	LB_NOUNROLL
	for (int iy = 1; iy <= col; iy++) {
		int is = col + iy;
		int is1 = LB_M + iy;
		LB_NOUNROLL
		for (int jy = 1; jy <= iy; jy++) {
			int js = col + jy;
			int js1 = LB_M + jy;
			wn[jy][iy] = wn1[iy][jy] / theta;
			wn[js][is] = wn1[is1][js1] * theta;
		}
		LB_NOUNROLL
		for (int jy = 1; jy <= iy - 1; jy++) {
			wn[jy][is] = -wn1[is1][jy];
		}
		LB_NOUNROLL
		for (int jy = iy; jy <= col; jy++) {
			wn[jy][is] = wn1[is1][jy];
		}
		wn[iy][iy] += s->sy[iy][iy];
	}
	/* first Cholesky: (1,1) block of WN */
	if (lb_dpofa(&wn[0][0], LB_M2 + 1, 0, col) != 0) {
		return -1;
	}
	/* then form L^-1(-L_a'+R_z') in the (1,2) block */
	const int col2 = 2 * col;
	LB_NOUNROLL
	for (int js = col + 1; js <= col2; js++) {
		/* column js of wn, rows 1..col, as the right-hand side of trans(T) x = b */
		double b[LB_M + 1];
		LB_NOUNROLL
		for (int i = 1; i <= col; i++) {
			b[i] = wn[i][js];
		}
		(void)lb_dtrsl(&wn[0][0], LB_M2 + 1, col, b, 11);
		LB_NOUNROLL
		for (int i = 1; i <= col; i++) {
			wn[i][js] = b[i];
		}
	}
	/* form S'AA'S*theta + (L^-1(-L_a'+R_z'))'(L^-1(-L_a'+R_z')) in the upper triangle of (2,2) */
	LB_NOUNROLL
	for (int is = col + 1; is <= col2; is++) {
		LB_NOUNROLL
		for (int js = is; js <= col2; js++) {
			double dot = 0.0;
			LB_NOUNROLL
			for (int i = 1; i <= col; i++) {
				dot += wn[i][is] * wn[i][js];
			}
			wn[is][js] += dot;
		}
	}
	/* Cholesky factorisation of the (2,2) block */
	if (lb_dpofa(&wn[0][0], LB_M2 + 1, col, col) != 0) {
		return -2;
	}
	return 0;
}

/* r = -Z'B(xcp - xk) - Z'g (subalgorithms.c cmprlb, :1305-1391); the problem is always constrained. */
LB_FN int lb_cmprlb(struct lb_state *s)
{
	const int col = s->col;
	LB_NOUNROLL
	for (int i = 1; i <= s->nfree; i++) {
		int k = s->index[i];
		s->r[i] = -s->theta * (s->z[k] - s->x[k]) - s->g[k];
	}
	/* p = wa[1..2m] receives M * c with c = wa[2m+1..4m] */
	if (lb_bmv(s, &s->wa[2 * LB_M], &s->wa[0]) != 0) {
		return -8;
	}
	int pointr = s->head;
	LB_NOUNROLL
	for (int j = 1; j <= col; j++) {
		double a1 = s->wa[j];
		double a2 = s->theta * s->wa[col + j];
		LB_NOUNROLL
		for (int i = 1; i <= s->nfree; i++) {
			int k = s->index[i];
			s->r[i] = s->r[i] + s->wy[k][pointr] * a1 + s->ws[k][pointr] * a2;
		}
		pointr = pointr % LB_M + 1;
	}
	return 0;
}

/* Subspace minimisation with the 2011 projection/backtracking refinement
 * (subalgorithms.c subsm, :1903-2228).  On entry z holds the Cauchy point, r the reduced gradient. */
LB_FN int lb_subsm(struct lb_state *s)
{
	const int col = s->col, nsub = s->nfree;
	const int *ind = s->index;
	double *x = s->z, *d = s->r, *xp = s->xp, *wv = s->wa;
	const double theta = s->theta;
	if (nsub <= 0) {
		return 0;
	}
	/* wv = W'Z d */
	int pointr = s->head;
	LB_NOUNROLL
	for (int i = 1; i <= col; i++) {
		double temp1 = 0.0, temp2 = 0.0;
		LB_NOUNROLL
		for (int j = 1; j <= nsub; j++) {
			int k = ind[j];
			temp1 += s->wy[k][pointr] * d[j];
			temp2 += s->ws[k][pointr] * d[j];
		}
		wv[i] = temp1;
		wv[col + i] = theta * temp2;
		pointr = pointr % LB_M + 1;
	}
	/* wv := K^-1 wv, K = LEL' stored as the upper-triangular factor in wn; the 2col x 2col system
	 * lives in the leading rows/columns of wn (its two blocks were packed contiguously by formk) */
	const int col2 = 2 * col;
	int info = lb_dtrsl(&s->wn[0][0], LB_M2 + 1, col2, wv, 11);
	if (info != 0) {
		return info;
	}
	LB_NOUNROLL
	for (int i = 1; i <= col; i++) {
		wv[i] = -wv[i];
	}
	info = lb_dtrsl(&s->wn[0][0], LB_M2 + 1, col2, wv, 1);
	if (info != 0) {
		return info;
	}
	/* d = (1/theta) d + (1/theta^2) Z'W wv */
	pointr = s->head;
	LB_NOUNROLL
	for (int jy = 1; jy <= col; jy++) {
		int js = col + jy;
		LB_NOUNROLL
		for (int i = 1; i <= nsub; i++) {
			int k = ind[i];
			d[i] = d[i] + s->wy[k][pointr] * wv[jy] / theta + s->ws[k][pointr] * wv[js];
		}
		pointr = pointr % LB_M + 1;
	}
	{
		double inv = 1.0 / theta;
		LB_NOUNROLL
		for (int i = 1; i <= nsub; i++) {
			d[i] = inv * d[i];
		}
	}
	/* projected Newton step */
	int iword = 0;
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		xp[i] = x[i];
	}
	LB_NOUNROLL
	for (int i = 1; i <= nsub; i++) {
		int k = ind[i];
		double dk = d[i];
		double xk = x[k];
		xk = lb_max(s->l[k], xk + dk);
		x[k] = lb_min(s->u[k], xk);
		if (x[k] == s->l[k] || x[k] == s->u[k]) {
			iword = 1;
		}
	}
	if (iword == 0) {
		return 0;
	}
	/* check sign of the directional derivative */
	double dd_p = 0.0;
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		dd_p += (x[i] - s->x[i]) * s->g[i];
	}
	if (dd_p > 0.0) {
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			x[i] = xp[i];
		}
		double alpha = 1.0;
		double temp1 = alpha;
		int ibd = 0;
		LB_NOUNROLL
		for (int i = 1; i <= nsub; i++) {
			int k = ind[i];
			double dk = d[i];
			if (dk < 0.0) {
				double temp2 = s->l[k] - x[k];
				if (temp2 >= 0.0) {
					temp1 = 0.0;
				} else if (dk * alpha < temp2) {
					temp1 = temp2 / dk;
				}
			} else if (dk > 0.0) {
				double temp2 = s->u[k] - x[k];
				if (temp2 <= 0.0) {
					temp1 = 0.0;
				} else if (dk * alpha > temp2) {
					temp1 = temp2 / dk;
				}
			}
			if (temp1 < alpha) {
				alpha = temp1;
				ibd = i;
			}
		}
		if (alpha < 1.0) {
			double dk = d[ibd];
			int k = ind[ibd];
			if (dk > 0.0) {
				x[k] = s->u[k];
				d[ibd] = 0.0;
			} else if (dk < 0.0) {
				x[k] = s->l[k];
				d[ibd] = 0.0;
			}
		}
		LB_NOUNROLL
		for (int i = 1; i <= nsub; i++) {
			int k = ind[i];
			x[k] += alpha * d[i];
		}
	}
	return 0;
}

/* Store the newest correction pair and refresh S'S, S'Y (subalgorithms.c matupd, :1393-1511). */
LB_FN void lb_matupd(struct lb_state *s, double rr, double dr)
{
	if (s->iupdat <= LB_M) {
		s->col = s->iupdat;
		s->itail = (s->head + s->iupdat - 2) % LB_M + 1;
	} else {
		s->itail = s->itail % LB_M + 1;
		s->head = s->head % LB_M + 1;
	}
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		s->ws[i][s->itail] = s->d[i];
		s->wy[i][s->itail] = s->r[i];
	}
	s->theta = rr / dr;
	const int col = s->col;
	if (s->iupdat > LB_M) {
		/* move old information */
		LB_NOUNROLL
		for (int j = 1; j <= col - 1; j++) {
			LB_NOUNROLL
			for (int i = 0; i < j; i++) {
				s->ss[1 + i][j] = s->ss[2 + i][j + 1];
			}
			LB_NOUNROLL
			for (int i = 0; i < col - j; i++) {
				s->sy[j + i][j] = s->sy[j + 1 + i][j + 1];
			}
		}
	}
	int pointr = s->head;
	LB_NOUNROLL
	for (int j = 1; j <= col - 1; j++) {
		double a = 0.0, b = 0.0;
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			a += s->d[i] * s->wy[i][pointr];
		}
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			b += s->ws[i][pointr] * s->d[i];
		}
		s->sy[col][j] = a;
		s->ss[j][col] = b;
		pointr = pointr % LB_M + 1;
	}
	if (s->stp == 1.0) {
		s->ss[col][col] = s->dtd;
	} else {
		s->ss[col][col] = s->stp * s->stp * s->dtd;
	}
	s->sy[col][col] = dr;
}

/* T = theta*S'S + L*D^-1*L', Cholesky-factored in place (subalgorithms.c formt, :920-974). */
LB_FN int lb_formt(struct lb_state *s)
{
	const int col = s->col;
	LB_NOUNROLL
	for (int j = 1; j <= col; j++) {
		s->wt[1][j] = s->theta * s->ss[1][j];
	}
	LB_NOUNROLL
	for (int i = 2; i <= col; i++) {
		LB_NOUNROLL
		for (int j = i; j <= col; j++) {
			int k1 = (i < j ? i : j) - 1;
			double ddum = 0.0;
			LB_NOUNROLL
			for (int k = 1; k <= k1; k++) {
				ddum += s->sy[i][k] * s->sy[j][k] / s->sy[k][k];
			}
			s->wt[i][j] = ddum + s->theta * s->ss[i][j];
		}
	}
	if (lb_dpofa(&s->wt[0][0], LB_M + 1, 0, col) != 0) {
		return -3;
	}
	return 0;
}

/* Safeguarded cubic/quadratic step of Moré & Thuente (linesearch.c dcstep, :485-763). */
LB_FN void lb_dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp,
                      double fp, double dp, int *brackt, double stpmin, double stpmax)
{
	double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
	sgnd = dp * (*dx / lb_abs(*dx));
	if (fp > *fx) {
		theta = (*fx - fp) * 3.0 / (*stp - *stx) + *dx + dp;
		s = lb_max(lb_max(lb_abs(theta), lb_abs(*dx)), lb_abs(dp));
		double d1 = theta / s;
		gamma = s * sqrt(d1 * d1 - *dx / s * (dp / s));
		if (*stp < *stx) {
			gamma = -gamma;
		}
		p = gamma - *dx + theta;
		q = gamma - *dx + gamma + dp;
		r = p / q;
		stpc = *stx + r * (*stp - *stx);
		stpq = *stx + *dx / ((*fx - fp) / (*stp - *stx) + *dx) / 2.0 * (*stp - *stx);
		if (lb_abs(stpc - *stx) < lb_abs(stpq - *stx)) {
			stpf = stpc;
		} else {
			stpf = stpc + (stpq - stpc) / 2.0;
		}
		*brackt = 1;
	} else if (sgnd < 0.0) {
		theta = (*fx - fp) * 3.0 / (*stp - *stx) + *dx + dp;
		s = lb_max(lb_max(lb_abs(theta), lb_abs(*dx)), lb_abs(dp));
		double d1 = theta / s;
		gamma = s * sqrt(d1 * d1 - *dx / s * (dp / s));
		if (*stp > *stx) {
			gamma = -gamma;
		}
		p = gamma - dp + theta;
		q = gamma - dp + gamma + *dx;
		r = p / q;
		stpc = *stp + r * (*stx - *stp);
		stpq = *stp + dp / (dp - *dx) * (*stx - *stp);
		if (lb_abs(stpc - *stp) > lb_abs(stpq - *stp)) {
			stpf = stpc;
		} else {
			stpf = stpq;
		}
		*brackt = 1;
	} else if (lb_abs(dp) < lb_abs(*dx)) {
		theta = (*fx - fp) * 3.0 / (*stp - *stx) + *dx + dp;
		s = lb_max(lb_max(lb_abs(theta), lb_abs(*dx)), lb_abs(dp));
		double d3 = theta / s;
		gamma = s * sqrt(lb_max(0.0, d3 * d3 - *dx / s * (dp / s)));
		if (*stp > *stx) {
			gamma = -gamma;
		}
		p = gamma - dp + theta;
		q = gamma + (*dx - dp) + gamma;
		r = p / q;
		if (r < 0.0 && gamma != 0.0) {
			stpc = *stp + r * (*stx - *stp);
		} else if (*stp > *stx) {
			stpc = stpmax;
		} else {
			stpc = stpmin;
		}
		stpq = *stp + dp / (dp - *dx) * (*stx - *stp);
		if (*brackt) {
			if (lb_abs(stpc - *stp) < lb_abs(stpq - *stp)) {
				stpf = stpc;
			} else {
				stpf = stpq;
			}
			if (*stp > *stx) {
				stpf = lb_min(*stp + (*sty - *stp) * 0.66, stpf);
			} else {
				stpf = lb_max(*stp + (*sty - *stp) * 0.66, stpf);
			}
		} else {
			if (lb_abs(stpc - *stp) > lb_abs(stpq - *stp)) {
				stpf = stpc;
			} else {
				stpf = stpq;
			}
			stpf = lb_min(stpmax, stpf);
			stpf = lb_max(stpmin, stpf);
		}
	} else {
		if (*brackt) {
			theta = (fp - *fy) * 3.0 / (*sty - *stp) + *dy + dp;
			s = lb_max(lb_max(lb_abs(theta), lb_abs(*dy)), lb_abs(dp));
			double d1 = theta / s;
			gamma = s * sqrt(d1 * d1 - *dy / s * (dp / s));
			if (*stp > *sty) {
				gamma = -gamma;
			}
			p = gamma - dp + theta;
			q = gamma - dp + gamma + *dy;
			r = p / q;
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

/* One reverse-communication turn of the Moré–Thuente search (linesearch.c dcsrch, :161-483).
 * The stpmin = 0 / ftol / gtol / xtol argument checks of the START branch cannot fire with the
 * constants above and stp = 1 <= stpmax, gd < 0 (checked by the caller), so they are omitted. */
LB_FN void lb_dcsrch(struct lb_state *s, double f, double g, double *stp, double stpmax)
{
	if (s->ls_task == LS_START) {
		s->brackt = 0;
		s->stage = 1;
		s->finit = f;
		s->ginit = g;
		s->gtest = LB_FTOL * s->ginit;
		s->width = stpmax - LB_STPMIN;
		s->width1 = s->width / 0.5;
		s->stx = 0.0;
		s->fx = s->finit;
		s->gx = s->ginit;
		s->sty = 0.0;
		s->fy = s->finit;
		s->gy = s->ginit;
		s->stmin = 0.0;
		s->stmax = *stp + *stp * 4.0;
		s->ls_task = LS_FG;
		return;
	}
	const double ftest = s->finit + *stp * s->gtest;
	if (s->stage == 1 && f <= ftest && g >= 0.0) {
		s->stage = 2;
	}
	int task = LS_FG;
	if (s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) {
		task = LS_WARNING;
	}
	if (s->brackt && s->stmax - s->stmin <= LB_XTOL * s->stmax) {
		task = LS_WARNING;
	}
	if (*stp == stpmax && f <= ftest && g <= s->gtest) {
		task = LS_WARNING;
	}
	if (*stp == LB_STPMIN && (f > ftest || g >= s->gtest)) {
		task = LS_WARNING;
	}
	if (f <= ftest && lb_abs(g) <= LB_GTOL * (-s->ginit)) {
		task = LS_CONVERGENCE;
	}
	if (task != LS_FG) {
		s->ls_task = task;
		return;
	}
	if (s->stage == 1 && f <= s->fx && f > ftest) {
		double fm = f - *stp * s->gtest;
		double fxm = s->fx - s->stx * s->gtest;
		double fym = s->fy - s->sty * s->gtest;
		double gm = g - s->gtest;
		double gxm = s->gx - s->gtest;
		double gym = s->gy - s->gtest;
		lb_dcstep(&s->stx, &fxm, &gxm, &s->sty, &fym, &gym, stp, fm, gm, &s->brackt, s->stmin, s->stmax);
		s->fx = fxm + s->stx * s->gtest;
		s->fy = fym + s->sty * s->gtest;
		s->gx = gxm + s->gtest;
		s->gy = gym + s->gtest;
	} else {
		lb_dcstep(&s->stx, &s->fx, &s->gx, &s->sty, &s->fy, &s->gy, stp, f, g, &s->brackt, s->stmin, s->stmax);
	}
	if (s->brackt) {
		if (lb_abs(s->sty - s->stx) >= s->width1 * 0.66) {
			*stp = s->stx + (s->sty - s->stx) * 0.5;
		}
		s->width1 = s->width;
		s->width = lb_abs(s->sty - s->stx);
	}
	if (s->brackt) {
		s->stmin = lb_min(s->stx, s->sty);
		s->stmax = lb_max(s->stx, s->sty);
	} else {
		s->stmin = *stp + (*stp - s->stx) * 1.1;
		s->stmax = *stp + (*stp - s->stx) * 4.0;
	}
	*stp = lb_max(*stp, LB_STPMIN);
	*stp = lb_min(*stp, stpmax);
	if ((s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) ||
	    (s->brackt && s->stmax - s->stmin <= LB_XTOL * s->stmax)) {
		*stp = s->stx;
	}
	s->ls_task = LS_FG;
}

/* ---------------------------------------------------------------------------------------------
 * Reverse-communication driver (lbfgsb.c mainlb, :327-1063, together with the driver loop of
 * src/min_saxs.c:229-243).  lb_step() advances the optimiser until it needs the objective at
 * s->x (returns LB_NEED_EVAL; the caller fills s->f, s->g[1..2] and calls again) or until it has
 * terminated (returns LB_DONE; s->x, s->f hold what the reference's driver reads back, s->nfgv the
 * number of objective evaluations).  Keeping the evaluation OUTSIDE the optimiser is what lets a
 * GPU warp run 32 independent fits in lock-step: the optimiser logic diverges, the expensive
 * objective is evaluated convergently by all lanes.
 */
enum lb_phase { LB_PH_INIT = 0, LB_PH_FIRST_EVAL, LB_PH_LINESEARCH, LB_PH_DONE,
                LB_PH_B_FIRST, LB_PH_B_ACCEPTED, LB_PH_B_NEW_ITER };
enum lb_status { LB_DONE = 0, LB_NEED_EVAL = 1, LB_NEED_B = 2 };

LB_FN void lb_begin(struct lb_state *s, double x1, double x2, double l1, double u1, double l2, double u2, double factr)
{
	s->x[1] = x1; s->x[2] = x2;
	s->l[1] = l1; s->l[2] = l2;
	s->u[1] = u1; s->u[2] = u2;
	s->g[1] = 0.0; s->g[2] = 0.0;
	s->f = 0.0;
	/* the reference zeroes its whole workspace before every fit (src/min_saxs.c:217-221) */
	LB_NOUNROLL
	for (int i = 0; i <= LB_N; i++) {
		LB_NOUNROLL
		for (int j = 0; j <= LB_M; j++) { s->ws[i][j] = 0.0; s->wy[i][j] = 0.0; }
		s->z[i] = s->r[i] = s->d[i] = s->t[i] = s->xp[i] = 0.0;
		s->index[i] = s->iwhere[i] = s->indx2[i] = 0;
	}
	LB_NOUNROLL
	for (int i = 0; i <= LB_M; i++) {
		LB_NOUNROLL
		for (int j = 0; j <= LB_M; j++) { s->sy[i][j] = 0.0; s->ss[i][j] = 0.0; s->wt[i][j] = 0.0; }
	}
	LB_NOUNROLL
	for (int i = 0; i <= LB_M2; i++) {
		LB_NOUNROLL
		for (int j = 0; j <= LB_M2; j++) { s->wn[i][j] = 0.0; s->wn1[i][j] = 0.0; }
	}
	LB_NOUNROLL
	for (int i = 0; i <= 8 * LB_M; i++) { s->wa[i] = 0.0; }
	s->brackt = 0; s->stage = 0; s->ls_task = LS_START;
	s->ginit = s->gtest = s->gx = s->gy = s->finit = s->fx = s->fy = 0.0;
	s->stx = s->sty = s->stmin = s->stmax = s->width = s->width1 = 0.0;

	lb_reset_memory(s);
	s->iback = 0; s->itail = 0; s->nact = 0; s->ileave = 0; s->nenter = 0;
	s->fold = 0.0; s->dnorm = 0.0; s->gd = 0.0; s->stpmx = 0.0; s->sbgnrm = 0.0;
	s->stp = 0.0; s->gdold = 0.0; s->dtd = 0.0; s->iter = 0; s->nfgv = 0; s->nseg = 0;
	s->nfree = LB_N; s->ifun = 0; s->wrk = 0;
	s->tol = factr * LB_EPSMCH;
	s->phase = LB_PH_INIT;
}

/* One turn of the line search with f, g at the current x (labels 666/556 of mainlb plus the
 * bookkeeping that follows lnsrlb's return, lbfgsb.c:871-915).  Outcomes:
 *   LB_PH_LINESEARCH   a new trial x was written, the objective is wanted there
 *   LB_PH_B_ACCEPTED   the search ended normally: iterate accepted
 *   LB_PH_B_NEW_ITER   the search failed with memory in use: memory reset, start a new iteration
 *   LB_PH_DONE         abnormal termination with empty memory: previous iterate restored */
LB_FN int lb_linesearch_turn(struct lb_state *s)
{
	int info_ls = 0;
	int search_over = 0;
	{
		double acc = 0.0;
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			acc += s->g[i] * s->d[i];
		}
		s->gd = acc;
	}
	if (s->ifun == 0) {
		s->gdold = s->gd;
		if (s->gd >= 0.0) {
			info_ls = -4; /* ascent direction in projection */
		}
	}
	if (info_ls == 0) {
		lb_dcsrch(s, s->f, s->gd, &s->stp, s->stpmx);
		if (s->ls_task == LS_FG) {
			++s->ifun;
			++s->nfgv;
			s->iback = s->ifun - 1;
			if (s->stp == 1.0) {
				LB_NOUNROLL
				for (int i = 1; i <= LB_N; i++) {
					s->x[i] = s->z[i];
				}
			} else {
				LB_NOUNROLL
				for (int i = 1; i <= LB_N; i++) {
					s->x[i] = s->stp * s->d[i] + s->t[i];
				}
			}
		} else {
			search_over = 1;
		}
	}
	if (info_ls != 0 || s->iback >= 20) {
		/* restore the previous iterate */
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			s->x[i] = s->t[i];
			s->g[i] = s->r[i];
		}
		s->f = s->fold;
		if (s->col == 0) {
			/* abnormal termination in the line search */
			if (info_ls == 0) {
				--s->nfgv;
				--s->ifun;
				--s->iback;
			}
			++s->iter;
			return LB_PH_DONE;
		}
		if (info_ls == 0) {
			--s->nfgv;
		}
		lb_reset_memory(s);
		return LB_PH_B_NEW_ITER;
	}
	return search_over ? LB_PH_B_ACCEPTED : LB_PH_LINESEARCH;
}

/* Part A — what every fit does right after an objective evaluation: cheap, and the same code for
 * (nearly) all lanes.  Returns LB_NEED_EVAL, LB_NEED_B (s->phase says where part B enters) or LB_DONE. */
LB_FN int lb_step_a(struct lb_state *s)
{
	if (s->phase == LB_PH_LINESEARCH) {
		s->phase = lb_linesearch_turn(s);
		if (s->phase == LB_PH_LINESEARCH) {
			return LB_NEED_EVAL;
		}
		return s->phase == LB_PH_DONE ? LB_DONE : LB_NEED_B;
	}
	if (s->phase == LB_PH_FIRST_EVAL) {
		s->phase = LB_PH_B_FIRST;
		return LB_NEED_B;
	}
	if (s->phase == LB_PH_INIT) {
		/* active(): project the start into the box, all variables boxed (subalgorithms.c:7-118) */
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			if (s->x[i] <= s->l[i]) {
				s->x[i] = s->l[i];
			} else if (s->x[i] >= s->u[i]) {
				s->x[i] = s->u[i];
			}
			s->iwhere[i] = (s->u[i] - s->l[i] <= 0.0) ? 3 : 0;
		}
		s->phase = LB_PH_FIRST_EVAL;
		return LB_NEED_EVAL;
	}
	return s->phase == LB_PH_DONE ? LB_DONE : LB_NEED_B;
}

/* Part B — the iteration boundary: convergence tests, BFGS update, generalised Cauchy point,
 * subspace minimisation, line-search set-up and its first turn.  Long and branchy; the GPU kernel
 * runs it on warps packed with exactly the fits that need it.  Returns LB_NEED_EVAL or LB_DONE. */
LB_FN int lb_step_b(struct lb_state *s, const double pgtol)
{
	if (s->phase == LB_PH_B_NEW_ITER) {
		goto new_iteration;
	}
	if (s->phase == LB_PH_B_ACCEPTED) {
		goto accepted;
	}
	if (s->phase != LB_PH_B_FIRST) {
		return s->phase == LB_PH_DONE ? LB_DONE : LB_NEED_EVAL;
	}

	s->nfgv = 1;
	s->sbgnrm = lb_projgr(s);
	if (s->sbgnrm <= pgtol) {
		goto finished;
	}

new_iteration: /* label 222 of mainlb */
	/* ---- generalised Cauchy point ---- */
	if (lb_cauchy(s) != 0) {
		lb_reset_memory(s);
		goto new_iteration;
	}
	lb_freev(s);
	s->nact = LB_N - s->nfree;
	/* ---- subspace minimisation ---- */
	if (s->nfree != 0 && s->col != 0) {
		int info = 0;
		if (s->wrk) {
			info = lb_formk(s);
		}
		if (info != 0) {
			lb_reset_memory(s);
			goto new_iteration;
		}
		info = lb_cmprlb(s);
		if (info == 0) {
			info = lb_subsm(s);
		}
		if (info != 0) {
			lb_reset_memory(s);
			goto new_iteration;
		}
	}
	/* ---- line search along d = z - x (linesearch.c lnsrlb, :5-159) ---- */
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		s->d[i] = s->z[i] - s->x[i];
	}
	{
		double acc = 0.0;
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			acc += s->d[i] * s->d[i];
		}
		s->dtd = acc;
	}
	s->dnorm = sqrt(s->dtd);
	s->stpmx = 1e10;
	if (s->iter == 0) {
		s->stpmx = 1.0;
	} else {
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			const double a1 = s->d[i];
			if (a1 < 0.0) {
				const double a2 = s->l[i] - s->x[i];
				if (a2 >= 0.0) {
					s->stpmx = 0.0;
				} else if (a1 * s->stpmx < a2) {
					s->stpmx = a2 / a1;
				}
			} else if (a1 > 0.0) {
				const double a2 = s->u[i] - s->x[i];
				if (a2 <= 0.0) {
					s->stpmx = 0.0;
				} else if (a1 * s->stpmx > a2) {
					s->stpmx = a2 / a1;
				}
			}
		}
	}
	s->stp = 1.0;
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		s->t[i] = s->x[i];
		s->r[i] = s->g[i];
	}
	s->fold = s->f;
	s->ifun = 0;
	s->iback = 0;
	s->ls_task = LS_START;

	/* first turn of the search: uses f, g at the current iterate, no new evaluation needed */
	s->phase = lb_linesearch_turn(s);
	if (s->phase == LB_PH_LINESEARCH) {
		return LB_NEED_EVAL;
	}
	if (s->phase == LB_PH_B_NEW_ITER) {
		goto new_iteration;
	}
	if (s->phase == LB_PH_DONE) {
		return LB_DONE;
	}
	/* LB_PH_B_ACCEPTED cannot follow a START turn (the search always asks for one evaluation) */

accepted: /* label 777 */
	++s->iter;
	s->sbgnrm = lb_projgr(s);
	if (s->sbgnrm <= pgtol) {
		goto finished;
	}
	{
		const double ddum = lb_max(lb_max(lb_abs(s->fold), lb_abs(s->f)), 1.0);
		if (s->fold - s->f <= s->tol * ddum) {
			goto finished;
		}
	}
	/* ---- BFGS update ---- */
	LB_NOUNROLL
	for (int i = 1; i <= LB_N; i++) {
		s->r[i] = s->g[i] - s->r[i];
	}
	{
		double rr = 0.0;
		LB_NOUNROLL
		for (int i = 1; i <= LB_N; i++) {
			rr += s->r[i] * s->r[i];
		}
		double dr, ddum2;
		if (s->stp == 1.0) {
			dr = s->gd - s->gdold;
			ddum2 = -s->gdold;
		} else {
			dr = (s->gd - s->gdold) * s->stp;
			LB_NOUNROLL
			for (int i = 1; i <= LB_N; i++) {
				s->d[i] = s->stp * s->d[i];
			}
			ddum2 = -s->gdold * s->stp;
		}
		if (dr <= LB_EPSMCH * ddum2) {
			s->updatd = 0; /* skip the update */
			goto new_iteration;
		}
		s->updatd = 1;
		++s->iupdat;
		lb_matupd(s, rr, dr);
		if (lb_formt(s) != 0) {
			lb_reset_memory(s);
		}
	}
	goto new_iteration;

finished:
	s->phase = LB_PH_DONE;
	return LB_DONE;
}

/* Serial composition of the two parts (host harness, single fits). */
LB_FN int lb_step(struct lb_state *s, const double pgtol)
{
	int r = lb_step_a(s);
	if (r == LB_NEED_B) {
		r = lb_step_b(s, pgtol);
	}
	return r;
}

#endif /* SXS_LBFGSB_N2M3_H */
