// pnp_math.h — small dense fp64 linear algebra shared by the oracle's PnP restatement.
// TEST INFRASTRUCTURE (oracle/): plain C++, no dependency on the product.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace orc_pnp {

// Eigen-decomposition of a symmetric n x n matrix (row-major, n <= 12): Householder reduction to tridiagonal
// form followed by the implicit-shift QL iteration (Wilkinson), every sum taken in ascending index order and
// the plane rotations normalised with sqrt(f*f + g*g) so that the CUDA kernels (rtabmap_b200/csrc/pnp_device.cuh,
// sym_eigen) can follow the same operation sequence.  This is what the PnP restatement uses wherever OpenCV takes
// the SVD of a small symmetric Gram matrix (MtM of EPnP, JtJ of the LM step, A^T A of the least-squares solves);
// jacobi_eigen_desc below is kept as an independent cross-check (tests/test_oracle_verify.py).
// On return: w[k] eigenvalues in DESCENDING order, vt row k = the matching unit eigenvector.
inline void sym_eigen_desc(const double * a_in, int n, double * w, double * vt)
{
	double a[144], z[144], d[12], e[12], v[12], p[12], wv[12];
	memcpy(a, a_in, sizeof(double) * n * n);
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n; ++j) z[i * n + j] = i == j ? 1.0 : 0.0;
	// ---- A = Q T Q^T, Q = H_0 H_1 ... accumulated in z
	for (int k = 0; k + 2 < n; ++k)
	{
		double sigma = 0;
		for (int i = k + 2; i < n; ++i) sigma += a[i * n + k] * a[i * n + k];
		if (sigma == 0.0) continue; // column already tridiagonal
		const double x0 = a[(k + 1) * n + k];
		const double nrm = std::sqrt(x0 * x0 + sigma);
		const double alpha = x0 > 0 ? -nrm : nrm;
		v[k + 1] = x0 - alpha;
		for (int i = k + 2; i < n; ++i) v[i] = a[i * n + k];
		const double beta = 2.0 / (v[k + 1] * v[k + 1] + sigma);
		for (int i = k + 1; i < n; ++i)
		{
			double s = 0;
			for (int j = k + 1; j < n; ++j) s += a[i * n + j] * v[j];
			p[i] = beta * s;
		}
		double vp = 0;
		for (int i = k + 1; i < n; ++i) vp += v[i] * p[i];
		const double K = 0.5 * beta * vp;
		for (int i = k + 1; i < n; ++i) wv[i] = p[i] - K * v[i];
		for (int i = k + 1; i < n; ++i)
			for (int j = k + 1; j < n; ++j) a[i * n + j] -= v[i] * wv[j] + wv[i] * v[j];
		a[(k + 1) * n + k] = alpha;
		a[k * n + k + 1] = alpha;
		for (int i = k + 2; i < n; ++i)
		{
			a[i * n + k] = 0.0;
			a[k * n + i] = 0.0;
		}
		for (int r = 0; r < n; ++r)
		{
			double s = 0;
			for (int j = k + 1; j < n; ++j) s += z[r * n + j] * v[j];
			s *= beta;
			for (int j = k + 1; j < n; ++j) z[r * n + j] -= s * v[j];
		}
	}
	for (int i = 0; i < n; ++i)
	{
		d[i] = a[i * n + i];
		e[i] = i + 1 < n ? a[(i + 1) * n + i] : 0.0;
	}
	// ---- implicit-shift QL on (d, e), rotations accumulated in the columns of z
	for (int l = 0; l < n; ++l)
	{
		for (int iter = 0; iter < 60; ++iter)
		{
			int m = l;
			for (; m + 1 < n; ++m)
				if (std::fabs(e[m]) <= 2.220446049250313e-16 * (std::fabs(d[m]) + std::fabs(d[m + 1]))) break;
			if (m == l) break;
			double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
			double r = std::sqrt(g * g + 1.0);
			g = d[m] - d[l] + e[l] / (g + (g >= 0 ? r : -r));
			double sn = 1.0, cs = 1.0, pp = 0.0;
			int i = m - 1;
			for (; i >= l; --i)
			{
				double f = sn * e[i];
				const double b = cs * e[i];
				r = std::sqrt(f * f + g * g);
				e[i + 1] = r;
				if (r == 0.0)
				{
					d[i + 1] -= pp;
					e[m] = 0.0;
					break;
				}
				sn = f / r;
				cs = g / r;
				g = d[i + 1] - pp;
				r = (d[i] - g) * sn + 2.0 * cs * b;
				pp = sn * r;
				d[i + 1] = g + pp;
				g = cs * r - b;
				for (int k = 0; k < n; ++k)
				{
					f = z[k * n + i + 1];
					z[k * n + i + 1] = sn * z[k * n + i] + cs * f;
					z[k * n + i] = cs * z[k * n + i] - sn * f;
				}
			}
			if (r == 0.0 && i >= l) continue;
			d[l] -= pp;
			e[l] = g;
			e[m] = 0.0;
		}
	}
	int order[12];
	for (int i = 0; i < n; ++i) order[i] = i;
	for (int i = 0; i < n - 1; ++i) // selection sort, descending, stable
	{
		int j = i;
		for (int k = i + 1; k < n; ++k)
			if (d[order[k]] > d[order[j]]) j = k;
		const int t = order[j];
		for (int k = j; k > i; --k) order[k] = order[k - 1];
		order[i] = t;
	}
	for (int k = 0; k < n; ++k)
	{
		w[k] = d[order[k]];
		for (int i = 0; i < n; ++i) vt[k * n + i] = z[i * n + order[k]];
	}
}

// Cyclic two-sided Jacobi eigen-decomposition of a symmetric n x n matrix (row-major, n <= 12).
// Rotations annihilate a_pq exactly; a pair is skipped once |a_pq| <= eps*sqrt(|a_pp a_qq|) and the
// iteration stops after a sweep without rotations (at most 30 sweeps).  The CUDA kernels use the same
// sequence of operations (rtabmap_b200/csrc/pnp_device.cuh, sym_eigen).
// On return: w[k] eigenvalues in DESCENDING order, vt row k = the matching unit eigenvector
// (the layout cvSVD(A, W, Ut, 0, CV_SVD_U_T) gives for a symmetric positive semi-definite A).
inline void jacobi_eigen_desc(const double * a_in, int n, double * w, double * vt)
{
	double a[144], v[144];
	memcpy(a, a_in, sizeof(double) * n * n);
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n; ++j) v[i * n + j] = i == j ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 30; ++sweep)
	{
		bool rotated = false;
		for (int p = 0; p < n - 1; ++p)
		{
			for (int q = p + 1; q < n; ++q)
			{
				const double apq = a[p * n + q];
				const double app = a[p * n + p], aqq = a[q * n + q];
				if (std::fabs(apq) <= 2.220446049250313e-16 * std::sqrt(std::fabs(app * aqq))) continue;
				rotated = true;
				const double theta = (aqq - app) / (2.0 * apq);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
				const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
				a[p * n + p] = app - t * apq;
				a[q * n + q] = aqq + t * apq;
				a[p * n + q] = 0.0;
				a[q * n + p] = 0.0;
				for (int k = 0; k < n; ++k)
				{
					if (k == p || k == q) continue;
					const double akp = a[k * n + p], akq = a[k * n + q];
					const double nkp = c * akp - s * akq, nkq = s * akp + c * akq;
					a[k * n + p] = nkp;
					a[p * n + k] = nkp;
					a[k * n + q] = nkq;
					a[q * n + k] = nkq;
				}
				for (int k = 0; k < n; ++k)
				{
					const double vkp = v[k * n + p], vkq = v[k * n + q];
					v[k * n + p] = c * vkp - s * vkq;
					v[k * n + q] = s * vkp + c * vkq;
				}
			}
		}
		if (!rotated) break;
	}
	int order[12];
	for (int i = 0; i < n; ++i) order[i] = i;
	for (int i = 0; i < n - 1; ++i) // selection sort, descending, stable
	{
		int j = i;
		for (int k = i + 1; k < n; ++k)
			if (a[order[k] * n + order[k]] > a[order[j] * n + order[j]]) j = k;
		const int t = order[j];
		for (int k = j; k > i; --k) order[k] = order[k - 1];
		order[i] = t;
	}
	for (int k = 0; k < n; ++k)
	{
		w[k] = a[order[k] * n + order[k]];
		for (int i = 0; i < n; ++i) vt[k * n + i] = v[i * n + order[k]];
	}
}

// One-sided (Hestenes) Jacobi SVD exactly as OpenCV's internal JacobiSVDImpl_ (modules/core/src/lapack.cpp),
// which cv::SVD / cvSVD use for matrices smaller than 25x25 even in LAPACK builds (hal_internal.cpp,
// HAL_SVD_SMALL_MATRIX_THRESH).  At is n x m row-major and holds A^T on entry (row i = column i of A);
// on return row i of At = i-th LEFT singular vector (unit), W = singular values descending, Vt (n x n,
// may be null) = V^T.  EPnP's control points depend on the SIGN of these vectors, which is why the
// iteration order, rotation formulas and the final selection sort are reproduced literally.
inline void jacobi_svd_opencv(double * At, double * W, double * Vt, int m, int n)
{
	const double eps = 2.220446049250313e-16 * 10, minval = 2.2250738585072014e-308;
	const int max_iter = std::max(m, 30);
	for (int i = 0; i < n; ++i)
	{
		double sd = 0;
		for (int k = 0; k < m; ++k) sd += At[i * m + k] * At[i * m + k];
		W[i] = sd;
		if (Vt)
		{
			for (int k = 0; k < n; ++k) Vt[i * n + k] = 0;
			Vt[i * n + i] = 1;
		}
	}
	for (int iter = 0; iter < max_iter; ++iter)
	{
		bool changed = false;
		for (int i = 0; i < n - 1; ++i)
			for (int j = i + 1; j < n; ++j)
			{
				double * Ai = At + i * m;
				double * Aj = At + j * m;
				double a = W[i], p = 0, b = W[j];
				for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
				if (std::fabs(p) <= eps * std::sqrt(a * b)) continue;
				p *= 2;
				const double beta = a - b, gamma = hypot(p, beta);
				double c, s;
				if (beta < 0)
				{
					const double delta = (gamma - beta) * 0.5;
					s = std::sqrt(delta / gamma);
					c = p / (gamma * s * 2);
				}
				else
				{
					c = std::sqrt((gamma + beta) / (gamma * 2));
					s = p / (gamma * c * 2);
				}
				a = b = 0;
				for (int k = 0; k < m; ++k)
				{
					const double t0 = c * Ai[k] + s * Aj[k];
					const double t1 = -s * Ai[k] + c * Aj[k];
					Ai[k] = t0;
					Aj[k] = t1;
					a += t0 * t0;
					b += t1 * t1;
				}
				W[i] = a;
				W[j] = b;
				changed = true;
				if (Vt)
				{
					double * Vi = Vt + i * n;
					double * Vj = Vt + j * n;
					for (int k = 0; k < n; ++k)
					{
						const double t0 = c * Vi[k] + s * Vj[k];
						const double t1 = -s * Vi[k] + c * Vj[k];
						Vi[k] = t0;
						Vj[k] = t1;
					}
				}
			}
		if (!changed) break;
	}
	for (int i = 0; i < n; ++i)
	{
		double sd = 0;
		for (int k = 0; k < m; ++k) sd += At[i * m + k] * At[i * m + k];
		W[i] = std::sqrt(sd);
	}
	for (int i = 0; i < n - 1; ++i)
	{
		int j = i;
		for (int k = i + 1; k < n; ++k)
			if (W[j] < W[k]) j = k;
		if (i != j)
		{
			std::swap(W[i], W[j]);
			for (int k = 0; k < m; ++k) std::swap(At[i * m + k], At[j * m + k]);
			if (Vt)
				for (int k = 0; k < n; ++k) std::swap(Vt[i * n + k], Vt[j * n + k]);
		}
	}
	for (int i = 0; i < n; ++i)
	{
		// OpenCV fills rows whose singular value is <= DBL_MIN with random orthogonal vectors; exact
		// rank deficiency does not occur on the noisy inputs of this path, so such rows are left as zeros.
		const double s = W[i] > minval ? 1 / W[i] : 0.;
		for (int k = 0; k < m; ++k) At[i * m + k] *= s;
	}
}

// Least-squares solution of A x = b (A is m x n, row-major, m >= n <= 6) through the normal
// equations' eigen-decomposition = the minimum-norm SVD solution cvSolve(..., CV_SVD) returns.
inline void solve_ls(const double * A, const double * b, int m, int n, double * x)
{
	double ata[36], atb[6], w[6], vt[36];
	for (int i = 0; i < n; ++i)
	{
		for (int j = 0; j < n; ++j)
		{
			double s = 0;
			for (int k = 0; k < m; ++k) s += A[k * n + i] * A[k * n + j];
			ata[i * n + j] = s;
		}
		double s = 0;
		for (int k = 0; k < m; ++k) s += A[k * n + i] * b[k];
		atb[i] = s;
	}
	sym_eigen_desc(ata, n, w, vt);
	for (int i = 0; i < n; ++i) x[i] = 0;
	const double tol = w[0] * 1e-14 * n;
	for (int k = 0; k < n; ++k)
	{
		if (w[k] <= tol) continue;
		double c = 0;
		for (int i = 0; i < n; ++i) c += vt[k * n + i] * atb[i];
		c /= w[k];
		for (int i = 0; i < n; ++i) x[i] += c * vt[k * n + i];
	}
}

// x = A^-1 b for a symmetric positive definite n x n matrix (n <= 6) by Cholesky, A = L L^T, sums in ascending index order.
// Returns false (x untouched) when a pivot is not safely positive (<= 1e-12 * the largest diagonal entry): the caller then
// falls back to the eigen-decomposition pseudo-inverse.  The damped normal equations of the LM step (diag * (1 + lambda)) are
// positive definite for every non-degenerate pose, where this gives the same step as cv::solve(DECOMP_SVD) to rounding.
// Same operation sequence as chol_solve in rtabmap_b200/csrc/pnp_device.cuh.
inline bool chol_solve(const double * A, const double * b, int n, double * x)
{
	double L[36], y[6];
	double dmax = 0;
	for (int i = 0; i < n; ++i) dmax = std::max(dmax, A[i * n + i]);
	for (int j = 0; j < n; ++j)
	{
		double d = A[j * n + j];
		for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
		if (!(d > 1e-12 * dmax)) return false;
		const double ljj = std::sqrt(d);
		L[j * n + j] = ljj;
		for (int i = j + 1; i < n; ++i)
		{
			double s = A[i * n + j];
			for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
			L[i * n + j] = s / ljj;
		}
	}
	for (int i = 0; i < n; ++i)
	{
		double s = b[i];
		for (int k = 0; k < i; ++k) s -= L[i * n + k] * y[k];
		y[i] = s / L[i * n + i];
	}
	for (int i = n - 1; i >= 0; --i)
	{
		double s = y[i];
		for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
		x[i] = s / L[i * n + i];
	}
	return true;
}

// Least squares of A x = b (A is m x n row-major, m >= n <= 6, full column rank) by Householder QR — what EPnP's
// Gauss-Newton refinement of the betas uses (OpenCV epnp.cpp, epnp::qr_solve on the 6x4 Jacobian).  A and b are
// destroyed.  Returns false (x untouched) when a column is exactly zero.  Same operation sequence as qr_solve in
// rtabmap_b200/csrc/pnp_device.cuh.
inline bool qr_solve_ls(double * A, double * b, int m, int n, double * x, double rel_tol = 0.0)
{
	double rdiag[6], v[12];
	double amax = 0;
	for (int k = 0; k < n; ++k)
	{
		double sigma = 0;
		for (int i = k; i < m; ++i) sigma += A[i * n + k] * A[i * n + k];
		if (sigma == 0.0) return false;
		const double akk = A[k * n + k];
		const double alpha = akk > 0 ? -std::sqrt(sigma) : std::sqrt(sigma);
		if (!(std::fabs(alpha) > rel_tol * amax)) return false; // numerically rank deficient: the caller uses the pseudo-inverse
		amax = std::max(amax, std::fabs(alpha));
		const double beta = 1.0 / (sigma - akk * alpha);
		v[k] = akk - alpha;
		for (int i = k + 1; i < m; ++i) v[i] = A[i * n + k];
		for (int j = k + 1; j < n; ++j)
		{
			double s = 0;
			for (int i = k; i < m; ++i) s += v[i] * A[i * n + j];
			s *= beta;
			for (int i = k; i < m; ++i) A[i * n + j] -= s * v[i];
		}
		double s = 0;
		for (int i = k; i < m; ++i) s += v[i] * b[i];
		s *= beta;
		for (int i = k; i < m; ++i) b[i] -= s * v[i];
		rdiag[k] = alpha;
	}
	for (int k = n - 1; k >= 0; --k)
	{
		double s = b[k];
		for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * x[j];
		x[k] = s / rdiag[k];
	}
	return true;
}

// cvSolve(A, b, x, CV_SVD) of the small systems of EPnP's find_betas_approx_*: Householder QR when the columns are safely
// independent (|R_kk| > 1e-8 * the largest |R_jj| so far), the eigen pseudo-inverse (minimum-norm solution) otherwise.
inline void ls_solve(const double * A, const double * b, int m, int n, double * x)
{
	double Ac[36], bc[6];
	memcpy(Ac, A, sizeof(double) * m * n);
	memcpy(bc, b, sizeof(double) * m);
	if (qr_solve_ls(Ac, bc, m, n, x, 1e-8)) return;
	solve_ls(A, b, m, n, x);
}

// SVD of a 3x3 matrix M = U diag(w) V^T (row-major U and V, columns are singular vectors).
inline void svd3(const double * M, double * U, double * w, double * V)
{
	double mtm[9], ew[3], vt[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
		{
			double s = 0;
			for (int k = 0; k < 3; ++k) s += M[k * 3 + i] * M[k * 3 + j];
			mtm[i * 3 + j] = s;
		}
	sym_eigen_desc(mtm, 3, ew, vt);
	for (int k = 0; k < 3; ++k)
	{
		w[k] = std::sqrt(std::max(ew[k], 0.0));
		for (int i = 0; i < 3; ++i) V[i * 3 + k] = vt[k * 3 + i];
	}
	// U columns = M v_k / w_k; complete the basis when a singular value vanishes
	for (int k = 0; k < 3; ++k)
	{
		double u[3];
		for (int i = 0; i < 3; ++i) u[i] = M[i * 3 + 0] * V[0 * 3 + k] + M[i * 3 + 1] * V[1 * 3 + k] + M[i * 3 + 2] * V[2 * 3 + k];
		double nrm = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
		if (nrm > 1e-12 * (w[0] + 1e-300))
			for (int i = 0; i < 3; ++i) U[i * 3 + k] = u[i] / nrm;
		else
		{
			// cross product of the two previous columns (k == 2) keeps U orthonormal
			const int a = (k + 1) % 3, b = (k + 2) % 3;
			U[0 * 3 + k] = U[1 * 3 + a] * U[2 * 3 + b] - U[2 * 3 + a] * U[1 * 3 + b];
			U[1 * 3 + k] = U[2 * 3 + a] * U[0 * 3 + b] - U[0 * 3 + a] * U[2 * 3 + b];
			U[2 * 3 + k] = U[0 * 3 + a] * U[1 * 3 + b] - U[1 * 3 + a] * U[0 * 3 + b];
		}
	}
}

inline void mat3_inv(const double * m, double * inv)
{
	const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
	const double id = 1.0 / det;
	inv[0] = (m[4] * m[8] - m[5] * m[7]) * id;
	inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
	inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
	inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
	inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
	inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
	inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
	inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
	inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// cv::Rodrigues, vector -> matrix, optional jacobian J[3][9] = dR(k)/dr(i)
inline void rodrigues_v2m(const double r[3], double R[9], double * J)
{
	const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
	if (theta < 2.220446049250313e-16)
	{
		for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
		if (J)
		{
			memset(J, 0, sizeof(double) * 27);
			J[5] = J[15] = J[19] = -1;
			J[7] = J[11] = J[21] = 1;
		}
		return;
	}
	const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, itheta = 1.0 / theta;
	const double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta;
	const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
	const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
	const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	for (int k = 0; k < 9; ++k) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
	if (J)
	{
		const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0, 0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
		const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
		for (int i = 0; i < 3; ++i)
		{
			const double ri = i == 0 ? rx : i == 1 ? ry : rz;
			const double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
			const double a3 = (c - s * itheta) * ri, a4 = s * itheta;
			for (int k = 0; k < 9; ++k) J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
		}
	}
}

// cv::Rodrigues, matrix -> vector (R is re-orthonormalised by SVD first, as OpenCV does)
inline void rodrigues_m2v(const double Rin[9], double r[3])
{
	double U[9], w[3], V[9], R[9];
	svd3(Rin, U, w, V);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) R[i * 3 + j] = U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
	double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
	const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
	double c = (R[0] + R[4] + R[8] - 1) * 0.5;
	c = c > 1. ? 1. : c < -1. ? -1. : c;
	const double theta = std::acos(c);
	if (s < 1e-5)
	{
		if (c > 0) { r[0] = r[1] = r[2] = 0; }
		else
		{
			double t;
			t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.));
			t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
			t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
			if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
			const double k = theta / std::sqrt(rx * rx + ry * ry + rz * rz);
			r[0] = rx * k; r[1] = ry * k; r[2] = rz * k;
		}
	}
	else
	{
		const double vth = 1 / (2 * s) * theta;
		r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
	}
}

} // namespace orc_pnp
