// oracle_verify.cpp — CPU restatement of the geometric-verification half of the hot path.
//
// THIS IS TEST INFRASTRUCTURE (see oracle.cpp header): only tests/, smoke() and bench.py's CPU legs
// may load it.  Restated from (paths under /root/reference/corelib/src):
//   pair matching           RegistrationVis.cpp:1482-1546  (temporary VWDictionary: addNewWords(from,1),
//                           update(), addNewWords(to,2); ids seen exactly once on each side are kept)
//   correspondence assembly util3d_motion_estimation.cpp:88-110 (ascending word id, finite 3D only)
//   RANSAC driver           opencv/solvepnp.cpp:112-211 (cv3::solvePnPRansac), :245-417
//                           (RANSACPointSetRegistrator::getSubset/findInliers/run), :213-233
//                           (RANSACUpdateNumIters), :63-101 (PnPRansacCallback)
//   refinement              util3d_motion_estimation.cpp:810-990 (computeReprojErrors, solvePnPRansac)
//   pose -> Transform       util3d_motion_estimation.cpp:121-154
// Third-party arithmetic that is NOT under /root/reference: OpenCV calib3d —
// cv::solvePnP(SOLVEPNP_EPNP) [EPnP, Lepetit/Moreno-Noguer/Fua 2009, OpenCV modules/calib3d/src/epnp.cpp],
// cv::solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess) [Levenberg-Marquardt on reprojection error,
// CvLevMarq, 20 iterations / FLT_EPSILON], cv::projectPoints, cv::Rodrigues, cv::RNG (MWC).  They are
// restated here from the published algorithms and pinned against the opencv-python 4.13 build in this
// image by tests/golden/make_pnp_golden.py -> tests/golden/pnp_golden.json.
#include "pnp_math.h"

#include <cfloat>
#include <cstdint>
#include <map>
#include <set>
#include <vector>

using namespace orc_pnp;

namespace {

int g_epnp_sign_mask = 0;

struct Cam
{
	double fu, fv, uc, vc;
};

// ---------------------------------------------------------------------------------- EPnP ----
struct Epnp
{
	Cam cam;
	int n = 0;
	std::vector<double> pws, us, alphas, pcs;
	double cws[4][3], ccs[4][3];

	void choose_control_points()
	{
		cws[0][0] = cws[0][1] = cws[0][2] = 0;
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < 3; ++j) cws[0][j] += pws[3 * i + j];
		for (int j = 0; j < 3; ++j) cws[0][j] /= n;
		double ptp[9] = {0};
		for (int i = 0; i < n; ++i)
		{
			double d[3];
			for (int j = 0; j < 3; ++j) d[j] = pws[3 * i + j] - cws[0][j];
			for (int a = 0; a < 3; ++a)
				for (int b = 0; b < 3; ++b) ptp[3 * a + b] += d[a] * d[b];
		}
		// cvSVD(&PW0tPW0, &DC, &UCt, 0, CV_SVD_MODIFY_A | CV_SVD_U_T): UCt rows = left singular vectors with
		// the signs OpenCV's internal Jacobi SVD produces (EPnP's answer on noisy samples depends on them)
		double dc[3], uct[9];
		for (int a = 0; a < 3; ++a)
			for (int b = 0; b < 3; ++b) uct[3 * a + b] = ptp[3 * b + a];
		jacobi_svd_opencv(uct, dc, nullptr, 3, 3);
		for (int r = 0; r < 3; ++r)
			if (g_epnp_sign_mask & (1 << r))
				for (int c = 0; c < 3; ++c) uct[3 * r + c] = -uct[3 * r + c];
		for (int i = 1; i < 4; ++i)
		{
			const double k = std::sqrt(std::max(dc[i - 1], 0.0) / n);
			for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * uct[3 * (i - 1) + j];
		}
	}

	void compute_barycentric_coordinates()
	{
		double cc[9], cci[9];
		for (int i = 0; i < 3; ++i)
			for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
		mat3_inv(cc, cci);
		alphas.assign(4 * n, 0.0);
		for (int i = 0; i < n; ++i)
		{
			const double * pi = &pws[3 * i];
			double * a = &alphas[4 * i];
			for (int j = 0; j < 3; ++j)
				a[1 + j] = cci[3 * j] * (pi[0] - cws[0][0]) + cci[3 * j + 1] * (pi[1] - cws[0][1]) + cci[3 * j + 2] * (pi[2] - cws[0][2]);
			a[0] = 1.0 - a[1] - a[2] - a[3];
		}
	}

	void compute_ccs(const double * betas, const double * ut)
	{
		for (int i = 0; i < 4; ++i) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0;
		for (int i = 0; i < 4; ++i)
		{
			const double * v = ut + 12 * (11 - i);
			for (int j = 0; j < 4; ++j)
				for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[3 * j + k];
		}
	}

	void compute_pcs()
	{
		pcs.assign(3 * n, 0.0);
		for (int i = 0; i < n; ++i)
		{
			const double * a = &alphas[4 * i];
			for (int j = 0; j < 3; ++j) pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
		}
	}

	void solve_for_sign()
	{
		if (pcs[2] < 0.0)
		{
			for (int i = 0; i < 4; ++i)
				for (int j = 0; j < 3; ++j) ccs[i][j] = -ccs[i][j];
			for (int i = 0; i < n; ++i)
				for (int j = 0; j < 3; ++j) pcs[3 * i + j] = -pcs[3 * i + j];
		}
	}

	void estimate_R_and_t(double R[3][3], double t[3])
	{
		double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < 3; ++j)
			{
				pc0[j] += pcs[3 * i + j];
				pw0[j] += pws[3 * i + j];
			}
		for (int j = 0; j < 3; ++j)
		{
			pc0[j] /= n;
			pw0[j] /= n;
		}
		double abt[9] = {0};
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < 3; ++j)
				for (int k = 0; k < 3; ++k) abt[3 * j + k] += (pcs[3 * i + j] - pc0[j]) * (pws[3 * i + k] - pw0[k]);
		double U[9], w[3], V[9];
		svd3(abt, U, w, V);
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j) R[i][j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
		const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] - R[0][2] * R[1][1] * R[2][0] -
		                   R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
		if (det < 0)
		{
			R[2][0] = -R[2][0];
			R[2][1] = -R[2][1];
			R[2][2] = -R[2][2];
		}
		for (int i = 0; i < 3; ++i) t[i] = pc0[i] - (R[i][0] * pw0[0] + R[i][1] * pw0[1] + R[i][2] * pw0[2]);
	}

	double reprojection_error(const double R[3][3], const double t[3])
	{
		double sum2 = 0.0;
		for (int i = 0; i < n; ++i)
		{
			const double * pw = &pws[3 * i];
			const double Xc = R[0][0] * pw[0] + R[0][1] * pw[1] + R[0][2] * pw[2] + t[0];
			const double Yc = R[1][0] * pw[0] + R[1][1] * pw[1] + R[1][2] * pw[2] + t[1];
			const double inv_Zc = 1.0 / (R[2][0] * pw[0] + R[2][1] * pw[1] + R[2][2] * pw[2] + t[2]);
			const double ue = cam.uc + cam.fu * Xc * inv_Zc, ve = cam.vc + cam.fv * Yc * inv_Zc;
			const double u = us[2 * i], v = us[2 * i + 1];
			sum2 += std::sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
		}
		return sum2 / n;
	}

	double compute_R_and_t(const double * ut, const double * betas, double R[3][3], double t[3])
	{
		compute_ccs(betas, ut);
		compute_pcs();
		solve_for_sign();
		estimate_R_and_t(R, t);
		return reprojection_error(R, t);
	}

	static void compute_L_6x10(const double * ut, double * l)
	{
		const double * v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
		double dv[4][6][3];
		for (int i = 0; i < 4; ++i)
		{
			int a = 0, b = 1;
			for (int j = 0; j < 6; ++j)
			{
				for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * a + k] - v[i][3 * b + k];
				if (++b > 3)
				{
					++a;
					b = a + 1;
				}
			}
		}
		auto dot = [](const double * x, const double * y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
		for (int i = 0; i < 6; ++i)
		{
			double * row = l + 10 * i;
			row[0] = dot(dv[0][i], dv[0][i]);
			row[1] = 2.0 * dot(dv[0][i], dv[1][i]);
			row[2] = dot(dv[1][i], dv[1][i]);
			row[3] = 2.0 * dot(dv[0][i], dv[2][i]);
			row[4] = 2.0 * dot(dv[1][i], dv[2][i]);
			row[5] = dot(dv[2][i], dv[2][i]);
			row[6] = 2.0 * dot(dv[0][i], dv[3][i]);
			row[7] = 2.0 * dot(dv[1][i], dv[3][i]);
			row[8] = 2.0 * dot(dv[2][i], dv[3][i]);
			row[9] = dot(dv[3][i], dv[3][i]);
		}
	}

	void compute_rho(double * rho)
	{
		auto d2 = [&](int a, int b) {
			return (cws[a][0] - cws[b][0]) * (cws[a][0] - cws[b][0]) + (cws[a][1] - cws[b][1]) * (cws[a][1] - cws[b][1]) +
			       (cws[a][2] - cws[b][2]) * (cws[a][2] - cws[b][2]);
		};
		rho[0] = d2(0, 1);
		rho[1] = d2(0, 2);
		rho[2] = d2(0, 3);
		rho[3] = d2(1, 2);
		rho[4] = d2(1, 3);
		rho[5] = d2(2, 3);
	}

	static void find_betas_approx_1(const double * l, const double * rho, double * betas)
	{
		double A[24], b4[4];
		for (int i = 0; i < 6; ++i)
		{
			A[4 * i] = l[10 * i];
			A[4 * i + 1] = l[10 * i + 1];
			A[4 * i + 2] = l[10 * i + 3];
			A[4 * i + 3] = l[10 * i + 6];
		}
		ls_solve(A, rho, 6, 4, b4);
		if (b4[0] < 0)
		{
			betas[0] = std::sqrt(-b4[0]);
			betas[1] = -b4[1] / betas[0];
			betas[2] = -b4[2] / betas[0];
			betas[3] = -b4[3] / betas[0];
		}
		else
		{
			betas[0] = std::sqrt(b4[0]);
			betas[1] = b4[1] / betas[0];
			betas[2] = b4[2] / betas[0];
			betas[3] = b4[3] / betas[0];
		}
	}

	static void find_betas_approx_2(const double * l, const double * rho, double * betas)
	{
		double A[18], b3[3];
		for (int i = 0; i < 6; ++i)
		{
			A[3 * i] = l[10 * i];
			A[3 * i + 1] = l[10 * i + 1];
			A[3 * i + 2] = l[10 * i + 2];
		}
		ls_solve(A, rho, 6, 3, b3);
		if (b3[0] < 0)
		{
			betas[0] = std::sqrt(-b3[0]);
			betas[1] = (b3[2] < 0) ? std::sqrt(-b3[2]) : 0.0;
		}
		else
		{
			betas[0] = std::sqrt(b3[0]);
			betas[1] = (b3[2] > 0) ? std::sqrt(b3[2]) : 0.0;
		}
		if (b3[1] < 0) betas[0] = -betas[0];
		betas[2] = 0.0;
		betas[3] = 0.0;
	}

	static void find_betas_approx_3(const double * l, const double * rho, double * betas)
	{
		double A[30], b5[5];
		for (int i = 0; i < 6; ++i)
			for (int j = 0; j < 5; ++j) A[5 * i + j] = l[10 * i + j];
		ls_solve(A, rho, 6, 5, b5);
		if (b5[0] < 0)
		{
			betas[0] = std::sqrt(-b5[0]);
			betas[1] = (b5[2] < 0) ? std::sqrt(-b5[2]) : 0.0;
		}
		else
		{
			betas[0] = std::sqrt(b5[0]);
			betas[1] = (b5[2] > 0) ? std::sqrt(b5[2]) : 0.0;
		}
		if (b5[1] < 0) betas[0] = -betas[0];
		betas[2] = b5[3] / betas[0];
		betas[3] = 0.0;
	}

	static void gauss_newton(const double * l, const double * rho, double * betas)
	{
		for (int k = 0; k < 5; ++k)
		{
			double A[24], b[6], x[4];
			for (int i = 0; i < 6; ++i)
			{
				const double * rl = l + 10 * i;
				double * ra = A + 4 * i;
				ra[0] = 2 * rl[0] * betas[0] + rl[1] * betas[1] + rl[3] * betas[2] + rl[6] * betas[3];
				ra[1] = rl[1] * betas[0] + 2 * rl[2] * betas[1] + rl[4] * betas[2] + rl[7] * betas[3];
				ra[2] = rl[3] * betas[0] + rl[4] * betas[1] + 2 * rl[5] * betas[2] + rl[8] * betas[3];
				ra[3] = rl[6] * betas[0] + rl[7] * betas[1] + rl[8] * betas[2] + 2 * rl[9] * betas[3];
				b[i] = rho[i] - (rl[0] * betas[0] * betas[0] + rl[1] * betas[0] * betas[1] + rl[2] * betas[1] * betas[1] +
				                 rl[3] * betas[0] * betas[2] + rl[4] * betas[1] * betas[2] + rl[5] * betas[2] * betas[2] +
				                 rl[6] * betas[0] * betas[3] + rl[7] * betas[1] * betas[3] + rl[8] * betas[2] * betas[3] +
				                 rl[9] * betas[3] * betas[3]);
			}
			if (!qr_solve_ls(A, b, 6, 4, x)) return;
			for (int i = 0; i < 4; ++i) betas[i] += x[i];
		}
	}

	// points: object xyz (double), image in PIXELS (already distortion-free)
	void compute_pose(double R[3][3], double t[3])
	{
		choose_control_points();
		compute_barycentric_coordinates();
		std::vector<double> M(2 * n * 12, 0.0);
		for (int i = 0; i < n; ++i)
		{
			const double * as = &alphas[4 * i];
			double * m1 = &M[(2 * i) * 12];
			double * m2 = m1 + 12;
			for (int k = 0; k < 4; ++k)
			{
				m1[3 * k] = as[k] * cam.fu;
				m1[3 * k + 1] = 0.0;
				m1[3 * k + 2] = as[k] * (cam.uc - us[2 * i]);
				m2[3 * k] = 0.0;
				m2[3 * k + 1] = as[k] * cam.fv;
				m2[3 * k + 2] = as[k] * (cam.vc - us[2 * i + 1]);
			}
		}
		double mtm[144], d[12], ut[144];
		for (int a = 0; a < 12; ++a)
			for (int b = 0; b < 12; ++b)
			{
				double s = 0;
				for (int r = 0; r < 2 * n; ++r) s += M[r * 12 + a] * M[r * 12 + b];
				mtm[a * 12 + b] = s;
			}
		sym_eigen_desc(mtm, 12, d, ut);
		double l[60], rho[6];
		compute_L_6x10(ut, l);
		compute_rho(rho);
		double Betas[4][4], rep[4], Rs[4][3][3], ts[4][3];
		find_betas_approx_1(l, rho, Betas[1]);
		gauss_newton(l, rho, Betas[1]);
		rep[1] = compute_R_and_t(ut, Betas[1], Rs[1], ts[1]);
		find_betas_approx_2(l, rho, Betas[2]);
		gauss_newton(l, rho, Betas[2]);
		rep[2] = compute_R_and_t(ut, Betas[2], Rs[2], ts[2]);
		find_betas_approx_3(l, rho, Betas[3]);
		gauss_newton(l, rho, Betas[3]);
		rep[3] = compute_R_and_t(ut, Betas[3], Rs[3], ts[3]);
		int N = 1;
		if (rep[2] < rep[1]) N = 2;
		if (rep[3] < rep[N]) N = 3;
		memcpy(R, Rs[N], sizeof(double) * 9);
		memcpy(t, ts[N], sizeof(double) * 3);
	}
};

// cv::solvePnP(SOLVEPNP_EPNP), no distortion: the pixel coordinates go through undistortPoints
// (normalised coordinates stored as float) and back, exactly as solvePnPGeneric does.
bool solve_pnp_epnp(const float * opts, const float * ipts, int n, const Cam & cam, double rvec[3], double tvec[3])
{
	Epnp e;
	e.cam = cam;
	e.n = n;
	e.pws.resize(3 * n);
	e.us.resize(2 * n);
	for (int i = 0; i < n; ++i)
	{
		for (int j = 0; j < 3; ++j) e.pws[3 * i + j] = opts[3 * i + j];
		const float xn = (float)(((double)ipts[2 * i] - cam.uc) / cam.fu);
		const float yn = (float)(((double)ipts[2 * i + 1] - cam.vc) / cam.fv);
		e.us[2 * i] = (double)xn * cam.fu + cam.uc;
		e.us[2 * i + 1] = (double)yn * cam.fv + cam.vc;
	}
	double R[3][3], t[3];
	e.compute_pose(R, t);
	for (int i = 0; i < 3; ++i)
		if (!std::isfinite(t[i])) return false;
	rodrigues_m2v(&R[0][0], rvec);
	memcpy(tvec, t, sizeof(double) * 3);
	return std::isfinite(rvec[0]) && std::isfinite(rvec[1]) && std::isfinite(rvec[2]);
}

// cv::projectPoints without distortion, optional jacobian rows (2 x 6 per point: d/drvec, d/dtvec)
void project(const float * opts, int n, const double rvec[3], const double tvec[3], const Cam & cam, double * uv, double * J)
{
	double R[9], dRdr[27];
	rodrigues_v2m(rvec, R, J ? dRdr : nullptr);
	for (int i = 0; i < n; ++i)
	{
		const double X = opts[3 * i], Y = opts[3 * i + 1], Z = opts[3 * i + 2];
		const double x = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
		const double y = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
		double z = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
		z = z ? 1. / z : 1;
		const double xn = x * z, yn = y * z;
		uv[2 * i] = xn * cam.fu + cam.uc;
		uv[2 * i + 1] = yn * cam.fv + cam.vc;
		if (J)
		{
			double * j0 = J + (2 * i) * 6;
			double * j1 = j0 + 6;
			// d/dt
			j0[3] = cam.fu * z;
			j0[4] = 0;
			j0[5] = -cam.fu * xn * z;
			j1[3] = 0;
			j1[4] = cam.fv * z;
			j1[5] = -cam.fv * yn * z;
			// d/dr
			for (int k = 0; k < 3; ++k)
			{
				const double * d = dRdr + 9 * k;
				const double dx = d[0] * X + d[1] * Y + d[2] * Z;
				const double dy = d[3] * X + d[4] * Y + d[5] * Z;
				const double dz = d[6] * X + d[7] * Y + d[8] * Z;
				j0[k] = cam.fu * (dx * z - xn * z * dz);
				j1[k] = cam.fv * (dy * z - yn * z * dz);
			}
		}
	}
}

// cv::solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess=true): CvLevMarq(6, 2n, 20 iters, FLT_EPSILON)
void solve_pnp_iterative_guess(const float * opts, const float * ipts, int n, const Cam & cam, double rvec[3], double tvec[3])
{
	double param[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]};
	double prev[6], JtJ[36], JtErr[6];
	std::vector<double> uv(2 * n), J(2 * n * 6), err(2 * n);
	auto eval = [&](const double * p, bool jac) {
		project(opts, n, p, p + 3, cam, uv.data(), jac ? J.data() : nullptr);
		double nrm = 0;
		for (int i = 0; i < 2 * n; ++i)
		{
			err[i] = uv[i] - (double)ipts[i];
			nrm += err[i] * err[i];
		}
		return std::sqrt(nrm);
	};
	auto step = [&](int lambdaLg10) {
		const double lambda = std::exp(lambdaLg10 * std::log(10.0));
		double A[36], w[6], vt[36], dx[6] = {0, 0, 0, 0, 0, 0};
		memcpy(A, JtJ, sizeof(A));
		for (int i = 0; i < 6; ++i) A[i * 6 + i] *= 1.0 + lambda;
		if (chol_solve(A, JtErr, 6, dx))
		{
			for (int i = 0; i < 6; ++i) param[i] = prev[i] - dx[i];
			return;
		}
		for (int i = 0; i < 6; ++i) dx[i] = 0;
		sym_eigen_desc(A, 6, w, vt);
		for (int k = 0; k < 6; ++k)
		{
			if (w[k] <= DBL_EPSILON * 6 * w[0]) continue;
			double c = 0;
			for (int i = 0; i < 6; ++i) c += vt[k * 6 + i] * JtErr[i];
			c /= w[k];
			for (int i = 0; i < 6; ++i) dx[i] += c * vt[k * 6 + i];
		}
		for (int i = 0; i < 6; ++i) param[i] = prev[i] - dx[i];
	};
	int lambdaLg10 = -3, iters = 0;
	const int max_iter = 20;
	const double eps = FLT_EPSILON;
	double prevErrNorm = DBL_MAX;
	for (;;)
	{
		// CALC_J
		double errNorm0 = eval(param, true);
		for (int a = 0; a < 6; ++a)
		{
			for (int b = 0; b < 6; ++b)
			{
				double s = 0;
				for (int r = 0; r < 2 * n; ++r) s += J[r * 6 + a] * J[r * 6 + b];
				JtJ[a * 6 + b] = s;
			}
			double s = 0;
			for (int r = 0; r < 2 * n; ++r) s += J[r * 6 + a] * err[r];
			JtErr[a] = s;
		}
		memcpy(prev, param, sizeof(prev));
		step(lambdaLg10);
		if (iters == 0) prevErrNorm = errNorm0;
		// CHECK_ERR
		bool done = false;
		for (;;)
		{
			const double errNorm = eval(param, false);
			if (errNorm > prevErrNorm)
			{
				if (++lambdaLg10 <= 16)
				{
					step(lambdaLg10);
					continue;
				}
			}
			lambdaLg10 = std::max(lambdaLg10 - 1, -16);
			double num = 0, den = 0;
			for (int i = 0; i < 6; ++i)
			{
				num += (param[i] - prev[i]) * (param[i] - prev[i]);
				den += prev[i] * prev[i];
			}
			if (++iters >= max_iter || std::sqrt(num) < eps * std::sqrt(den)) done = true; // cvNorm(param, prevParam, CV_RELATIVE_L2) < eps
			prevErrNorm = errNorm;
			break;
		}
		if (done) break;
	}
	for (int i = 0; i < 3; ++i)
	{
		rvec[i] = param[i];
		tvec[i] = param[3 + i];
	}
}

// cv::RNG (multiply-with-carry), state (uint64)-1 as in RANSACPointSetRegistrator::run
struct CvRng
{
	uint64_t state = 0xFFFFFFFFFFFFFFFFull;
	unsigned next()
	{
		state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
		return (unsigned)state;
	}
	int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

int ransac_update_num_iters(double p, double ep, int modelPoints, int maxIters)
{
	p = std::max(p, 0.);
	p = std::min(p, 1.);
	ep = std::max(ep, 0.);
	ep = std::min(ep, 1.);
	double num = std::max(1. - p, DBL_MIN);
	double denom = 1. - std::pow(1. - ep, modelPoints);
	if (denom < DBL_MIN) return 0;
	num = std::log(num);
	denom = std::log(denom);
	return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)std::nearbyint(num / denom); // cvRound
}

// PnPRansacCallback::computeError + findInliers: err = (float)norm(ipt - (Point2f)proj), inlier iff err <= (float)(thr*thr)
int find_inliers(const float * opts, const float * ipts, int n, const Cam & cam, const double rvec[3], const double tvec[3], double thr,
                 std::vector<unsigned char> & mask)
{
	std::vector<double> uv(2 * n);
	project(opts, n, rvec, tvec, cam, uv.data(), nullptr);
	const float t = (float)(thr * thr);
	int nz = 0;
	mask.resize(n);
	for (int i = 0; i < n; ++i)
	{
		const float dx = ipts[2 * i] - (float)uv[2 * i], dy = ipts[2 * i + 1] - (float)uv[2 * i + 1];
		const float e = (float)std::sqrt((double)dx * dx + (double)dy * dy);
		const int f = e <= t;
		mask[i] = (unsigned char)f;
		nz += f;
	}
	return nz;
}

struct PnpResult
{
	bool ok = false;
	double rvec[3] = {0, 0, 0}, tvec[3] = {0, 0, 0};
	std::vector<int> inliers;
	int iterations_run = 0;
};

// cv3::solvePnPRansac + util3d::solvePnPRansac refinement
PnpResult pnp_ransac(const float * opts, const float * ipts, int n, const Cam & cam, int iterations, float reproj, int min_inliers,
                     int refine_iterations, float refine_sigma, const double * guess_rt)
{
	PnpResult res;
	if (guess_rt)
	{
		memcpy(res.rvec, guess_rt, 3 * sizeof(double));
		memcpy(res.tvec, guess_rt + 3, 3 * sizeof(double));
	}
	if (min_inliers < 4) min_inliers = 4;
	const int model_points = 6; // npoints == 4 would use P3P: not reachable with Vis/MinInliers >= 6
	if (n < model_points) return res;
	const double confidence = 0.99;
	int niters = std::max(iterations, 1);
	CvRng rng;
	std::vector<unsigned char> mask, bestMask;
	double bestR[3] = {0, 0, 0}, bestT[3] = {0, 0, 0};
	double curR[3] = {res.rvec[0], res.rvec[1], res.rvec[2]}, curT[3] = {res.tvec[0], res.tvec[1], res.tvec[2]};
	int maxGood = 0;
	if (n == model_points)
	{
		if (!solve_pnp_epnp(opts, ipts, n, cam, curR, curT)) return res;
		memcpy(bestR, curR, sizeof(bestR));
		memcpy(bestT, curT, sizeof(bestT));
		bestMask.assign(n, 1);
		maxGood = n;
	}
	else
	{
		int iter;
		for (iter = 0; iter < niters; ++iter)
		{
			int idx[6];
			for (int i = 0; i < model_points;)
			{
				int idx_i;
				for (;;)
				{
					idx_i = idx[i] = rng.uniform(0, n);
					int j;
					for (j = 0; j < i; ++j)
						if (idx_i == idx[j]) break;
					if (j == i) break;
				}
				++i;
			}
			float so[18], si[12];
			for (int i = 0; i < 6; ++i)
			{
				memcpy(so + 3 * i, opts + 3 * idx[i], 3 * sizeof(float));
				memcpy(si + 2 * i, ipts + 2 * idx[i], 2 * sizeof(float));
			}
			if (!solve_pnp_epnp(so, si, 6, cam, curR, curT)) continue;
			const int good = find_inliers(opts, ipts, n, cam, curR, curT, reproj, mask);
			if (good > std::max(maxGood, model_points - 1))
			{
				std::swap(mask, bestMask);
				memcpy(bestR, curR, sizeof(bestR));
				memcpy(bestT, curT, sizeof(bestT));
				maxGood = good;
				niters = ransac_update_num_iters(confidence, (double)(n - good) / n, model_points, niters);
			}
		}
		res.iterations_run = iter;
	}
	if (maxGood <= 0) return res; // rvec/tvec keep the guess (solvepnp.cpp:183-192)
	// final solvePnP on the inliers is computed by the reference but its result is discarded: the
	// returned pose is the best minimal-sample model (solvepnp.cpp:196-197)
	memcpy(res.rvec, bestR, sizeof(bestR));
	memcpy(res.tvec, bestT, sizeof(bestT));
	for (int i = 0; i < n; ++i)
		if (bestMask[i]) res.inliers.push_back(i);
	res.ok = true;

	// util3d::solvePnPRansac refinement (util3d_motion_estimation.cpp:882-989)
	if ((int)res.inliers.size() >= min_inliers && refine_iterations > 0)
	{
		const float inlierThreshold = reproj;
		float error_threshold = inlierThreshold;
		int refine_it = 0;
		bool inlier_changed = false;
		std::vector<int> new_inliers, prev_inliers = res.inliers;
		std::vector<size_t> inliers_sizes;
		double mr[3], mt[3];
		memcpy(mr, res.rvec, sizeof(mr));
		memcpy(mt, res.tvec, sizeof(mt));
		do
		{
			std::vector<float> oi(3 * prev_inliers.size()), ii(2 * prev_inliers.size());
			for (size_t i = 0; i < prev_inliers.size(); ++i)
			{
				memcpy(&oi[3 * i], opts + 3 * prev_inliers[i], 3 * sizeof(float));
				memcpy(&ii[2 * i], ipts + 2 * prev_inliers[i], 2 * sizeof(float));
			}
			solve_pnp_iterative_guess(oi.data(), ii.data(), (int)prev_inliers.size(), cam, mr, mt);
			inliers_sizes.push_back(prev_inliers.size());
			// computeReprojErrors: e = (float)norm(ipt - proj) <= error_threshold (NOT squared here)
			std::vector<double> uv(2 * n);
			project(opts, n, mr, mt, cam, uv.data(), nullptr);
			new_inliers.assign(n, 0);
			std::vector<float> err(n);
			int oi_ = 0;
			for (int i = 0; i < n; ++i)
			{
				const float dx = ipts[2 * i] - (float)uv[2 * i], dy = ipts[2 * i + 1] - (float)uv[2 * i + 1];
				const float e = (float)std::sqrt((double)dx * dx + (double)dy * dy);
				if (e <= error_threshold)
				{
					new_inliers[oi_] = i;
					err[oi_++] = e;
				}
			}
			new_inliers.resize(oi_);
			err.resize(oi_);
			if ((int)new_inliers.size() < min_inliers)
			{
				++refine_it;
				if (refine_it >= refine_iterations) break;
				continue;
			}
			float m = 0;
			for (float v : err) m += v;
			m /= err.size();
			float variance = 0;
			if (err.size() > 1)
			{
				float sum = 0;
				for (float v : err) sum += (v - m) * (v - m);
				variance = sum / (err.size() - 1);
			}
			error_threshold = std::min(inlierThreshold, refine_sigma * float(std::sqrt(variance)));
			inlier_changed = false;
			std::swap(prev_inliers, new_inliers);
			if (new_inliers.size() != prev_inliers.size())
			{
				if ((int)inliers_sizes.size() >= min_inliers)
				{
					if (inliers_sizes[inliers_sizes.size() - 1] == inliers_sizes[inliers_sizes.size() - 3] &&
					    inliers_sizes[inliers_sizes.size() - 2] == inliers_sizes[inliers_sizes.size() - 4])
						break;
				}
				inlier_changed = true;
				continue;
			}
			for (size_t i = 0; i < prev_inliers.size(); ++i)
			{
				if (prev_inliers[i] != new_inliers[i])
				{
					inlier_changed = true;
					break;
				}
			}
		} while (inlier_changed && ++refine_it < refine_iterations);
		std::swap(res.inliers, new_inliers);
		memcpy(res.rvec, mr, sizeof(mr));
		memcpy(res.tvec, mt, sizeof(mt));
	}
	return res;
}

} // namespace

// entry points of oracle.cpp (the restated VWDictionary) used for the temporary matching dictionary
extern "C" {
void * orc_create(int desc_type, int dim, int incremental, float nndr, int cmp_new);
void orc_destroy(void * h);
void orc_update(void * h);
int orc_add_new_words(void * h, const void * desc, int n, int sig_id, int * out_ids);
}

extern "C" {

void orcv_set_sign_mask(int m) { g_epnp_sign_mask = m; }

// symmetric eigen-decomposition (method 0: Householder + QL, the one the restatement uses; 1: cyclic Jacobi cross-check)
void orcv_sym_eigen(const double * a, int n, int method, double * w, double * vt)
{
	if (method == 0) sym_eigen_desc(a, n, w, vt);
	else jacobi_eigen_desc(a, n, w, vt);
}

// RegistrationVis global matching (RegistrationVis.cpp:1482-1546) with Vis/CorNNType in {0,3}: a temporary
// incremental VWDictionary quantises the FROM descriptors (signature 1), is updated, then quantises the TO
// descriptors (signature 2).  from_ids[n_from], to_ids[n_to].
void orcv_match_pair(int desc_type, int dim, const void * desc_from, int n_from, const void * desc_to, int n_to, float nndr,
                     int * from_ids, int * to_ids)
{
	void * d = orc_create(desc_type, dim, 1, nndr, 1);
	if (n_from) orc_add_new_words(d, desc_from, n_from, 1, from_ids);
	if (n_to)
	{
		orc_update(d);
		orc_add_new_words(d, desc_to, n_to, 2, to_ids);
	}
	orc_destroy(d);
}

// Memory::computeTransform -> RegistrationVis (global matching, no guess) -> util3d::estimateMotion3DTo2D for one
// pair, single camera, identity localTransform, no distortion.  xyz_from[n_from*3] (NaN = no depth),
// uv_to[n_to*2].  Outputs: matches (word ids, ascending), inlier word ids, rvec/tvec of the PnP pose and
// transform[12] = (localTransform * pnp)^-1 as 3x4 float.  Returns 1 when inliers >= min_inliers.
int orcv_verify_pair(int desc_type, int dim, const void * desc_from, const float * xyz_from, int n_from, const void * desc_to,
                     const float * uv_to, int n_to, const double K[4], float nndr, int min_inliers, int iterations, float reproj,
                     int refine_iterations, int * match_ids, int * n_matches, int * inlier_ids, int * n_inliers, double rvec[3],
                     double tvec[3], float transform[12])
{
	std::vector<int> fid(std::max(n_from, 1)), tid(std::max(n_to, 1));
	orcv_match_pair(desc_type, dim, desc_from, n_from, desc_to, n_to, nndr, fid.data(), tid.data());
	std::multiset<int> fset(fid.begin(), fid.begin() + n_from), tset(tid.begin(), tid.begin() + n_to);
	std::map<int, int> wordsFrom, wordsTo; // unique word id -> descriptor index
	for (int i = 0; i < n_from; ++i)
		if (fset.count(fid[i]) == 1) wordsFrom[fid[i]] = i;
	for (int i = 0; i < n_to; ++i)
		if (tset.count(tid[i]) == 1) wordsTo[tid[i]] = i;
	std::vector<float> op, ip;
	std::vector<int> matches;
	for (auto & kv : wordsTo) // uKeys(words2B): ascending id
	{
		auto it = wordsFrom.find(kv.first);
		if (it == wordsFrom.end()) continue;
		const float * p = xyz_from + 3 * it->second;
		if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
		op.insert(op.end(), p, p + 3);
		ip.push_back(uv_to[2 * kv.second]);
		ip.push_back(uv_to[2 * kv.second + 1]);
		matches.push_back(kv.first);
	}
	*n_matches = (int)matches.size();
	for (size_t i = 0; i < matches.size(); ++i) match_ids[i] = matches[i];
	*n_inliers = 0;
	rvec[0] = rvec[1] = rvec[2] = tvec[0] = tvec[1] = tvec[2] = 0;
	for (int i = 0; i < 12; ++i) transform[i] = 0;
	if ((int)matches.size() < min_inliers) return 0;
	Cam cam{K[0], K[1], K[2], K[3]};
	const double guess[6] = {0, 0, 0, 0, 0, 0}; // identity guess (Memory::computeTransform passes a null guess -> identity)
	PnpResult r = pnp_ransac(op.data(), ip.data(), (int)matches.size(), cam, iterations, reproj, min_inliers, refine_iterations, 3.0f, guess);
	memcpy(rvec, r.rvec, sizeof(r.rvec));
	memcpy(tvec, r.tvec, sizeof(r.tvec));
	*n_inliers = (int)r.inliers.size();
	for (size_t i = 0; i < r.inliers.size(); ++i) inlier_ids[i] = matches[r.inliers[i]];
	if ((int)r.inliers.size() < min_inliers) return 0;
	double R[9];
	rodrigues_v2m(r.rvec, R, nullptr);
	// Transform pnp(R|t) as float, then the rigid inverse (localTransform = identity)
	float Rf[9], tf[3];
	for (int i = 0; i < 9; ++i) Rf[i] = (float)R[i];
	for (int i = 0; i < 3; ++i) tf[i] = (float)r.tvec[i];
	for (int i = 0; i < 3; ++i)
	{
		for (int j = 0; j < 3; ++j) transform[4 * i + j] = Rf[3 * j + i];
		transform[4 * i + 3] = -(Rf[0 + i] * tf[0] + Rf[3 + i] * tf[1] + Rf[6 + i] * tf[2]);
	}
	return 1;
}

// The covariance block of util3d::estimateMotion3DTo2D (util3d_motion_estimation.cpp:156-266) for one accepted pose, single camera,
// identity localTransform.  obj[n*3] / img[n*2]: the correspondences given to PnP; obj_to[n*3]: the 3-D point of camera B for each
// correspondence (words3B; NaN = none) or NULL when the TO signature has no 3-D words; inliers[n_inl] index the correspondences.
// T[12] = transform (float 3x4) as returned; rvec/tvec the PnP pose.  cov[36] row-major.  Returns 0 when Vis/PnPMaxVariance rejects
// the transform (covariance reset to identity), else 1.  Float arithmetic follows the reference statement by statement
// (util3d::transformPoint, util3d_transforms.cpp:211-220; util3d::projectDepthTo3DRay, util3d.cpp:246-264; uNormSquared, UMath.h:601-605;
// pcl::getAngle3D is third party (PCL common: acos of the clamped dot product of the normalised vectors)).
int orcv_covariance(const float * obj, const float * img, const float * obj_to, int n, const int * inliers, int n_inl, const float T[12],
                    const double rvec[3], const double tvec[3], const double K[4], int img_w, int img_h, int var_median_ratio,
                    float max_variance, int split_linear, double cov[36])
{
	for (int i = 0; i < 36; ++i) cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
	if (n_inl <= 0) return 1;
	auto tp = [](const float * M, float x, float y, float z, float out[3]) {
		out[0] = M[0] * x + M[1] * y + M[2] * z + M[3];
		out[1] = M[4] * x + M[5] * y + M[6] * z + M[7];
		out[2] = M[8] * x + M[9] * y + M[10] * z + M[11];
	};
	if (obj_to || img_w != 0 || img_h != 0)
	{
		// transformCameraFrameInv = (transform * localTransform)^-1: the rigid inverse of T
		float P[12];
		for (int i = 0; i < 3; ++i)
		{
			for (int j = 0; j < 3; ++j) P[4 * i + j] = T[4 * j + i];
			P[4 * i + 3] = -(T[0 + i] * T[3] + T[4 + i] * T[7] + T[8 + i] * T[11]);
		}
		std::vector<float> eD(n_inl), eA(n_inl), eX(n_inl), eY(n_inl), eZ(n_inl);
		for (int k = 0; k < n_inl; ++k)
		{
			const int i = inliers[k];
			const float ox = obj[3 * i], oy = obj[3 * i + 1], oz = obj[3 * i + 2];
			float np_[3];
			const float * b = obj_to ? obj_to + 3 * i : nullptr;
			if (b && std::isfinite(b[0]) && std::isfinite(b[1]) && std::isfinite(b[2])) tp(T, b[0], b[1], b[2], np_);
			else
			{
				float cb[3];
				tp(P, ox, oy, oz, cb);
				float cx = (float)K[2], cy = (float)K[3];
				const float fx = (float)K[0], fy = (float)K[1];
				cx = cx > 0.0f ? cx : float(img_w / 2) - 0.5f;
				cy = cy > 0.0f ? cy : float(img_h / 2) - 0.5f;
				const float rx = (img[2 * i] - cx) / fx, ry = (img[2 * i + 1] - cy) / fy;
				const double sc = cb[2] * 1.1; // float * double literal
				const float px = (float)(rx * sc), py = (float)(ry * sc), pz = (float)(1.0f * sc);
				tp(T, px, py, pz, np_);
			}
			const float dx = ox - np_[0], dy = oy - np_[1], dz = oz - np_[2];
			eD[k] = dx * dx + dy * dy + dz * dz;
			const double ex = dx, ey = dy, ez = dz;
			eX[k] = (float)(ex * ex);
			eY[k] = (float)(ey * ey);
			eZ[k] = (float)(ez * ez);
			const float n1 = std::sqrt(ox * ox + oz * oz + oy * oy), n2 = std::sqrt(np_[0] * np_[0] + np_[2] * np_[2] + np_[1] * np_[1]);
			const float ax = ox / n1, ay = oy / n1, az = oz / n1, bx = np_[0] / n2, by = np_[1] / n2, bz = np_[2] / n2;
			double rad = (double)(ax * bx + az * bz + ay * by);
			rad = rad < -1.0 ? -1.0 : (rad > 1.0 ? 1.0 : rad);
			eA[k] = (float)std::acos(rad);
		}
		const int kth = n_inl / var_median_ratio;
		std::sort(eD.begin(), eD.end());
		std::sort(eA.begin(), eA.end());
		double lin = 2.1981 * (double)eD[kth];
		const double ang = 2.1981 * (double)eA[kth];
		for (int i = 0; i < 3; ++i) cov[7 * i] *= lin;
		for (int i = 3; i < 6; ++i) cov[7 * i] *= ang;
		if (split_linear)
		{
			std::sort(eX.begin(), eX.end());
			std::sort(eY.begin(), eY.end());
			std::sort(eZ.begin(), eZ.end());
			const double mx = 2.1981 * (double)eX[kth], my = 2.1981 * (double)eY[kth], mz = 2.1981 * (double)eZ[kth];
			cov[0] = mx;
			cov[7] = my;
			cov[14] = mz;
			lin = std::max(mx, std::max(my, mz));
		}
		if (max_variance > 0 && lin > max_variance)
		{
			for (int i = 0; i < 36; ++i) cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
			return 0;
		}
		return 1;
	}
	// rms of the inliers' reprojection errors (:253-266)
	Cam cam{K[0], K[1], K[2], K[3]};
	std::vector<double> uvp(2 * n);
	project(obj, n, rvec, tvec, cam, uvp.data(), nullptr);
	float err = 0.0f;
	for (int k = 0; k < n_inl; ++k)
	{
		const int i = inliers[k];
		const float ex = img[2 * i] - (float)uvp[2 * i], ey = img[2 * i + 1] - (float)uvp[2 * i + 1];
		err += ex * ex + ey * ey;
	}
	const double sc = std::sqrt(err / float(n_inl));
	for (int i = 0; i < 6; ++i) cov[7 * i] *= sc;
	return 1;
}

static double g_cov_lin_eps = 0.00000001, g_cov_ang_eps = 0.00000003; // Registration::COVARIANCE_LINEAR / ANGULAR_EPSILON (Registration.cpp:36-37)
static void covariance_floor(double cov[36])
{
	// Registration::computeTransformationMod (Registration.cpp:239-250)
	for (int i = 0; i < 3; ++i)
		if (cov[7 * i] <= g_cov_lin_eps) cov[7 * i] = g_cov_lin_eps;
	for (int i = 3; i < 6; ++i)
		if (cov[7 * i] <= g_cov_ang_eps) cov[7 * i] = g_cov_ang_eps;
}

// orcv_verify_pair + the covariance: xyz_to[n_to*3] may be NULL.  cov[36].
int orcv_verify_pair_cov(int desc_type, int dim, const void * desc_from, const float * xyz_from, int n_from, const void * desc_to,
                         const float * uv_to, const float * xyz_to, int n_to, const double K[4], float nndr, int min_inliers, int iterations,
                         float reproj, int refine_iterations, int img_w, int img_h, int var_median_ratio, float max_variance, int split_linear,
                         int * match_ids, int * n_matches, int * inlier_ids, int * n_inliers, double rvec[3], double tvec[3], float transform[12],
                         double cov[36])
{
	for (int i = 0; i < 36; ++i) cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
	int ok = orcv_verify_pair(desc_type, dim, desc_from, xyz_from, n_from, desc_to, uv_to, n_to, K, nndr, min_inliers, iterations, reproj,
	                          refine_iterations, match_ids, n_matches, inlier_ids, n_inliers, rvec, tvec, transform);
	if (!ok) return 0;
	// rebuild the correspondence arrays in match order (match ids are the temporary dictionary's word ids)
	std::vector<int> fid(std::max(n_from, 1)), tid(std::max(n_to, 1));
	orcv_match_pair(desc_type, dim, desc_from, n_from, desc_to, n_to, nndr, fid.data(), tid.data());
	std::map<int, int> f_of, t_of;
	for (int i = 0; i < n_from; ++i) f_of[fid[i]] = i;
	for (int i = 0; i < n_to; ++i) t_of[tid[i]] = i;
	const int nm = *n_matches, ni = *n_inliers;
	std::vector<float> obj(3 * nm), img(2 * nm), objt(3 * nm);
	std::map<int, int> pos;
	for (int m = 0; m < nm; ++m)
	{
		const int id = match_ids[m], fi = f_of[id], ti = t_of[id];
		pos[id] = m;
		memcpy(&obj[3 * m], xyz_from + 3 * fi, 3 * sizeof(float));
		memcpy(&img[2 * m], uv_to + 2 * ti, 2 * sizeof(float));
		if (xyz_to) memcpy(&objt[3 * m], xyz_to + 3 * ti, 3 * sizeof(float));
	}
	std::vector<int> inl(ni);
	for (int k = 0; k < ni; ++k) inl[k] = pos[inlier_ids[k]];
	ok = orcv_covariance(obj.data(), img.data(), xyz_to ? objt.data() : nullptr, nm, inl.data(), ni, transform, rvec, tvec, K, img_w, img_h,
	                     var_median_ratio, max_variance, split_linear, cov);
	if (!ok)
		for (int i = 0; i < 12; ++i) transform[i] = 0; // transform.setNull()
	covariance_floor(cov);
	return ok;
}

// ---- second registration pass (Reg/RepeatOnce, Registration.cpp:221-229) ------------------------------------------------------
// RegistrationVis::computeTransformationImpl WITH a guess, default branch Vis/CorGuessMatchToProjection = false
// (RegistrationVis.cpp:1017-1070 projection, :1225-1370 matching): the 3-D points of FROM are projected into TO's image with the guess
// (cv::projectPoints, no distortion), the TO keypoints within Vis/CorGuessWinSize pixels of each projection are its candidates, the
// candidate set is ranked by cv::BFMatcher::knnMatch(k = 2) + the strict Vis/CorNNDR test (a single candidate is taken as it is), every TO
// keypoint goes to the first FROM point (ascending index) that selects it; ids = FROM indices.
// The reference finds the candidates with a randomised kd-tree limited to 32 checks (rtflann::KDTreeIndexParams(), SearchParams(32, 0,
// false)): approximate and unordered, i.e. not a deterministic function of the inputs.  This restatement — and the
// CUDA kernel it checks — use the EXACT radius search (squared distance < r^2 as RadiusResultSet::addPoint, rtflann/util/result_set.h:475-479)
// with candidates in ascending TO index: a superset of what the kd-tree can return, same rule otherwise.  Parity for this pass is
// therefore oracle-vs-CUDA only ("unpinned").
// sel_to[n_from]: TO index matched to FROM i (-1 none).  Returns the number of projected FROM points.
int orcv_guess_match(int desc_type, int dim, const void * desc_from, const float * xyz_from, int n_from, const void * desc_to, const float * uv_to,
                     int n_to, const double K[4], const double rvec[3], const double tvec[3], int img_w, int img_h, float win, float nndr, int * sel_to)
{
	Cam cam{K[0], K[1], K[2], K[3]};
	std::vector<double> uvp(2 * (size_t)std::max(n_from, 1));
	// projectPoints on the float points; NaN points give NaN pixels and fail the bounds test
	project(xyz_from, n_from, rvec, tvec, cam, uvp.data(), nullptr);
	double R[9];
	rodrigues_v2m(rvec, R, nullptr);
	const size_t rb = desc_type == 0 ? (size_t)dim : (size_t)dim * 4;
	std::vector<int> claim(std::max(n_to, 1), -1);
	int n_proj = 0;
	for (int i = 0; i < n_from; ++i) sel_to[i] = -1;
	for (int i = 0; i < n_from; ++i)
	{
		const float px = (float)uvp[2 * i], py = (float)uvp[2 * i + 1];
		const float * X = xyz_from + 3 * i;
		// util3d::transformPoint(kptsFrom3D[i], guessCameraRef).z with the float Transform of the guess
		const float zc = (float)R[6] * X[0] + (float)R[7] * X[1] + (float)R[8] * X[2] + (float)tvec[2];
		const bool inb = std::isfinite(px) && !(px < 0.0f) && !(px >= float(img_w - 1)) && std::isfinite(py) && !(py < 0.0f) && !(py >= float(img_h - 1));
		if (!(inb && zc > 0.0f)) continue;
		++n_proj;
		if (!(std::isfinite(X[0]) && std::isfinite(X[1]) && std::isfinite(X[2]))) continue;
		// candidates: exact radius search, ascending TO index; k = 2 brute force among them
		const float r2 = win * win;
		int cnt = 0, b1 = -1;
		float d1 = INFINITY, d2 = INFINITY;
		const uint8_t * qd = (const uint8_t *)desc_from + rb * i;
		for (int j = 0; j < n_to; ++j)
		{
			float dist = 0.0f;
			const float dx = px - uv_to[2 * j], dy = py - uv_to[2 * j + 1];
			dist += dx * dx;
			dist += dy * dy;
			if (!(dist < r2)) continue;
			++cnt;
			float d;
			const uint8_t * td = (const uint8_t *)desc_to + rb * j;
			if (desc_type == 0)
			{
				int h = 0;
				for (int k = 0; k < dim; ++k) h += __builtin_popcount((unsigned)(qd[k] ^ td[k]));
				d = (float)h;
			}
			else
			{
				const float * a = (const float *)qd;
				const float * b = (const float *)td;
				float acc = 0.0f;
				for (int k = 0; k < dim; k += 4)
				{
					const float e0 = a[k] - b[k], e1 = a[k + 1] - b[k + 1], e2 = a[k + 2] - b[k + 2], e3 = a[k + 3] - b[k + 3];
					acc += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
				}
				d = acc;
			}
			if (d < d1)
			{
				d2 = d1;
				d1 = d;
				b1 = j;
			}
			else if (d < d2) d2 = d;
		}
		int m = -1;
		if (cnt >= 2)
		{
			if (d1 < nndr * d2) m = b1;
		}
		else if (cnt == 1) m = b1;
		if (m >= 0 && claim[m] < 0)
		{
			claim[m] = i;
			sel_to[i] = m;
		}
	}
	return n_proj;
}


// Memory::computeTransform as Registration::computeTransformationMod runs it by default: pass 1 without a guess (global matching), then
// — Reg/RepeatOnce, only when pass 1 succeeded — pass 2 with pass 1's transform as the guess; its outputs replace pass 1's.
int orcv_verify_pair_repeat(int desc_type, int dim, const void * desc_from, const float * xyz_from, int n_from, const void * desc_to,
                            const float * uv_to, const float * xyz_to, int n_to, const double K[4], float nndr, int min_inliers, int iterations,
                            float reproj, int refine_iterations, int img_w, int img_h, int var_median_ratio, float max_variance, int split_linear,
                            int repeat_once, float guess_win, int * match_ids, int * n_matches, int * inlier_ids, int * n_inliers, double rvec[3],
                            double tvec[3], float transform[12], double cov[36], int * second_pass_ran)
{
	*second_pass_ran = 0;
	int ok = orcv_verify_pair_cov(desc_type, dim, desc_from, xyz_from, n_from, desc_to, uv_to, xyz_to, n_to, K, nndr, min_inliers, iterations, reproj,
	                              refine_iterations, img_w, img_h, var_median_ratio, max_variance, split_linear, match_ids, n_matches, inlier_ids,
	                              n_inliers, rvec, tvec, transform, cov);
	if (ok && repeat_once && guess_win > 0 && img_w > 0 && img_h > 0)
	{
		*second_pass_ran = 1;
		std::vector<int> sel(std::max(n_from, 1));
		orcv_guess_match(desc_type, dim, desc_from, xyz_from, n_from, desc_to, uv_to, n_to, K, rvec, tvec, img_w, img_h, guess_win, nndr, sel.data());
		std::vector<float> obj, img, objt;
		std::vector<int> ids;
		for (int i = 0; i < n_from; ++i)
		{
			if (sel[i] < 0) continue;
			const float * p = xyz_from + 3 * i;
			obj.insert(obj.end(), p, p + 3);
			img.push_back(uv_to[2 * sel[i]]);
			img.push_back(uv_to[2 * sel[i] + 1]);
			if (xyz_to) objt.insert(objt.end(), xyz_to + 3 * sel[i], xyz_to + 3 * sel[i] + 3);
			ids.push_back(i);
		}
		const int nm = (int)ids.size();
		*n_matches = nm;
		for (int m = 0; m < nm; ++m) match_ids[m] = ids[m];
		*n_inliers = 0;
		for (int i = 0; i < 12; ++i) transform[i] = 0;
		for (int i = 0; i < 36; ++i) cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
		ok = 0;
		if (nm >= min_inliers)
		{
			Cam cam{K[0], K[1], K[2], K[3]};
			const double guess[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]};
			PnpResult r = pnp_ransac(obj.data(), img.data(), nm, cam, iterations, reproj, min_inliers, refine_iterations, 3.0f, guess);
			memcpy(rvec, r.rvec, sizeof(r.rvec));
			memcpy(tvec, r.tvec, sizeof(r.tvec));
			*n_inliers = (int)r.inliers.size();
			for (size_t i = 0; i < r.inliers.size(); ++i) inlier_ids[i] = ids[r.inliers[i]];
			if ((int)r.inliers.size() >= min_inliers)
			{
				double R[9];
				rodrigues_v2m(r.rvec, R, nullptr);
				float Rf[9], tf[3];
				for (int i = 0; i < 9; ++i) Rf[i] = (float)R[i];
				for (int i = 0; i < 3; ++i) tf[i] = (float)r.tvec[i];
				for (int i = 0; i < 3; ++i)
				{
					for (int j = 0; j < 3; ++j) transform[4 * i + j] = Rf[3 * j + i];
					transform[4 * i + 3] = -(Rf[0 + i] * tf[0] + Rf[3 + i] * tf[1] + Rf[6 + i] * tf[2]);
				}
				ok = orcv_covariance(obj.data(), img.data(), xyz_to ? objt.data() : nullptr, nm, r.inliers.data(), (int)r.inliers.size(), transform, rvec, tvec,
				                     K, img_w, img_h, var_median_ratio, max_variance, split_linear, cov);
				if (!ok)
					for (int i = 0; i < 12; ++i) transform[i] = 0;
			}
		}
	}
	covariance_floor(cov);
	return ok;
}

// the restated OpenCV pieces, exposed one by one so tests can pin each against cv2
int orcv_solve_pnp_epnp(const float * opts, const float * ipts, int n, const double K[4], double rvec[3], double tvec[3])
{
	Cam cam{K[0], K[1], K[2], K[3]};
	return solve_pnp_epnp(opts, ipts, n, cam, rvec, tvec) ? 1 : 0;
}
void orcv_solve_pnp_iterative(const float * opts, const float * ipts, int n, const double K[4], double rvec[3], double tvec[3])
{
	Cam cam{K[0], K[1], K[2], K[3]};
	solve_pnp_iterative_guess(opts, ipts, n, cam, rvec, tvec);
}
void orcv_project(const float * opts, int n, const double K[4], const double rvec[3], const double tvec[3], double * uv)
{
	Cam cam{K[0], K[1], K[2], K[3]};
	project(opts, n, rvec, tvec, cam, uv, nullptr);
}
void orcv_rodrigues(const double rvec[3], double R[9]) { rodrigues_v2m(rvec, R, nullptr); }
void orcv_rodrigues_inv(const double R[9], double rvec[3]) { rodrigues_m2v(R, rvec); }
void orcv_rng_draws(int count, int n, unsigned * out)
{
	CvRng rng;
	for (int i = 0; i < count; ++i) out[i] = (unsigned)rng.uniform(0, n);
}

// util3d::solvePnPRansac (incl. cv3::solvePnPRansac): returns 1 if RANSAC found a model; inliers_out[n]
int orcv_pnp_ransac(const float * opts, const float * ipts, int n, const double K[4], int iterations, float reproj, int min_inliers,
                    int refine_iterations, float refine_sigma, const double * guess_rt, double rvec[3], double tvec[3],
                    int * inliers_out, int * n_inliers, int * iterations_run)
{
	Cam cam{K[0], K[1], K[2], K[3]};
	PnpResult r = pnp_ransac(opts, ipts, n, cam, iterations, reproj, min_inliers, refine_iterations, refine_sigma, guess_rt);
	memcpy(rvec, r.rvec, sizeof(r.rvec));
	memcpy(tvec, r.tvec, sizeof(r.tvec));
	*n_inliers = (int)r.inliers.size();
	for (size_t i = 0; i < r.inliers.size(); ++i) inliers_out[i] = r.inliers[i];
	if (iterations_run) *iterations_run = r.iterations_run;
	return r.ok ? 1 : 0;
}

} // extern "C"
