/* TEST INFRASTRUCTURE (see oracle/__init__.py): float64 truth for rows a1-a5 of SURVEY.md section 8 in plain C,
 * multi-threaded with OpenMP so that the FULL benchmark shape (200 704 points x 512 cameras) can be checked in
 * seconds.  Same closed form as oracle/sh_cov.py::sh_basis_closed_form_f64 (an independent, trig-free restatement
 * of the reference):
 *
 *   ray d = cam - pt;  cos(theta) = y/r, sin(theta) = rho/r, cos(phi) = z/rho, sin(phi) = x/rho, rho = sqrt(x^2+z^2)
 *   (reference utility/CustomGeometry.py:27-45 with theta = pi/2 - elev, networks/SconeVis.py:232-234)
 *   P_m^m = (-1)^m (2m-1)!! sin^m(theta);  P_l^m = ((2l-1) x P_{l-1}^m - (l+m-1) P_{l-2}^m) / (l-m)
 *   (reference utility/spherical_harmonics.py:67-108)
 *   Y_lm = N_lm P_l^|m| {cos(m phi) | 1 | sin(|m| phi)},  N_l0 = sqrt((2l+1)/4pi), N_lm = N_l0 sqrt(2 (l-|m|)!/(l+|m|)!)
 *   (reference utility/spherical_harmonics.py:111-156), column k = l*l + l + m
 *   out[b,c,p] = act(sum_k Y_k H[b,p,k]);  coverage[b,c] = mean_p  (reference networks/SconeVis.py:164-252)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's parity check call this; it is never part of the product path.
 */
#include <math.h>
#include <stddef.h>

#define N_DEGREE 8
#define N_HARM 64

static double g_norm[N_DEGREE][N_DEGREE]; /* N_lm, m >= 0 */
static double g_pmm[N_DEGREE];            /* (-1)^m (2m-1)!! */
static int g_ready = 0;

static void init_tables(void)
{
    if (g_ready) return;
    for (int m = 0; m < N_DEGREE; ++m) {
        double df = 1.0;
        for (int n = 2 * m - 1; n > 1; n -= 2) df *= n;
        g_pmm[m] = ((m & 1) ? -1.0 : 1.0) * df;
    }
    for (int l = 0; l < N_DEGREE; ++l)
        for (int m = 0; m <= l; ++m) {
            double n = sqrt((2 * l + 1) / (4.0 * M_PI));
            if (m) {
                double rising = 1.0; /* (l-m+1)(l-m+2)...(l+m) = (l+m)!/(l-m)! */
                for (int v = l - m + 1; v <= l + m; ++v) rising *= v;
                n *= sqrt(2.0 / rising);
            }
            g_norm[l][m] = n;
        }
    g_ready = 1;
}

/* z = sum_k Y_k(d) h[k] in float64; h are the fp32 coefficients of one point */
static double project(const double dx, const double dy, const double dz, const float *h)
{
    const double rho2 = dx * dx + dz * dz;
    const double r = sqrt(rho2 + dy * dy);
    const double rho = sqrt(rho2);
    const double ct = dy / r, st = rho / r;
    const double cp = dz / rho, sp = dx / rho;
    double cm[N_DEGREE], sm[N_DEGREE];
    cm[0] = 1.0;
    sm[0] = 0.0;
    for (int m = 1; m < N_DEGREE; ++m) {
        cm[m] = cm[m - 1] * cp - sm[m - 1] * sp;
        sm[m] = sm[m - 1] * cp + cm[m - 1] * sp;
    }
    double z = 0.0, stm = 1.0;
    for (int m = 0; m < N_DEGREE; ++m) {
        const double pmm = g_pmm[m] * stm;
        double prev2 = 0.0, prev1 = pmm;
        for (int l = m; l < N_DEGREE; ++l) {
            double p;
            if (l == m) p = pmm;
            else if (l == m + 1) p = (2 * m + 1) * ct * pmm;
            else p = ((2 * l - 1) * ct * prev1 - (l + m - 1) * prev2) / (l - m);
            if (l > m) {
                prev2 = prev1;
                prev1 = p;
            }
            const double np_ = g_norm[l][m] * p;
            const int k = l * l + l;
            if (m == 0) z += np_ * (double)h[k];
            else z += np_ * (cm[m] * (double)h[k + m] + sm[m] * (double)h[k - m]);
        }
        stm *= st;
    }
    return z;
}

static double activate(const double z, const int use_sigmoid)
{
    return use_sigmoid ? 1.0 / (1.0 + exp(-z)) : (z > 0.0 ? z : 0.0);
}

/* (B,P,pts_dim), (B,P,64), (B,C,3) fp32 -> coverage (B,C) float64 (mean over the P points) */
int mac_oracle_coverage_f64(const float *pts, int pts_dim, const float *harm, const float *cams, int B, int P, int C,
                            int use_sigmoid, double *out)
{
    if (!pts || !harm || !cams || !out || B <= 0 || P <= 0 || C <= 0 || pts_dim < 3) return -1;
    init_tables();
#pragma omp parallel for schedule(dynamic, 1)
    for (long bc = 0; bc < (long)B * C; ++bc) {
        const int b = (int)(bc / C);
        const float *cam = cams + (size_t)bc * 3;
        const double cx = cam[0], cy = cam[1], cz = cam[2];
        double sum = 0.0, comp = 0.0; /* Kahan: the sum of 200 k terms stays exact to ~1e-16 relative */
        for (int p = 0; p < P; ++p) {
            const float *pt = pts + ((size_t)b * P + p) * pts_dim;
            const double v = activate(project(cx - pt[0], cy - pt[1], cz - pt[2], harm + ((size_t)b * P + p) * N_HARM),
                                      use_sigmoid);
            const double y = v - comp, t = sum + y;
            comp = (t - sum) - y;
            sum = t;
        }
        out[bc] = sum / P;
    }
    return 0;
}

/* same inputs -> per-point values (B,C,P) float64 */
int mac_oracle_visibility_f64(const float *pts, int pts_dim, const float *harm, const float *cams, int B, int P, int C,
                              int use_sigmoid, double *out)
{
    if (!pts || !harm || !cams || !out || B <= 0 || P <= 0 || C <= 0 || pts_dim < 3) return -1;
    init_tables();
#pragma omp parallel for schedule(dynamic, 1)
    for (long bc = 0; bc < (long)B * C; ++bc) {
        const int b = (int)(bc / C);
        const float *cam = cams + (size_t)bc * 3;
        const double cx = cam[0], cy = cam[1], cz = cam[2];
        for (int p = 0; p < P; ++p) {
            const float *pt = pts + ((size_t)b * P + p) * pts_dim;
            out[(size_t)bc * P + p] = activate(
                project(cx - pt[0], cy - pt[1], cz - pt[2], harm + ((size_t)b * P + p) * N_HARM), use_sigmoid);
        }
    }
    return 0;
}
