/*
 * CPU oracle in plain C (TEST INFRASTRUCTURE ONLY -- never linked into or called by the product).
 *
 * Scalar FP64 restatement of the FLORIS 3.5 GCH steady-state solve that ifpen/wfcrl-env calls from
 * wfcrl/interface.py:564 (calculate_wake), :623 (get_turbine_powers), :629-648 (load proxies, local wind
 * measurements), configured by wfcrl/simulators/floris/inputs/template/case.yaml:14-16,27-39,41-60,84-89.
 * FLORIS itself is an un-vendored dependency (requirements.txt:8) that is absent here; the algorithm follows
 * SURVEY.md Appendix A/B and is validated against oracle/floris_oracle.py (numpy, pinned to the reference's
 * notebook vector examples/demo.ipynb:137-138) by tests/test_oracle.py.  Parity status: see floris_oracle.py.
 *
 * Purpose: a checker fast enough for batched parity tests at T=80 and thousands of envs.
 * Build: make -C oracle   (-> oracle/_build/liboracle.so)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define G 3
#define NP 9
#define NTAB 51

static const double WS_TAB[NTAB] = {
    0.0, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5, 6.0, 6.5, 7.0, 7.5, 8.0, 8.5, 9.0, 9.5, 10.0, 10.5, 11.0, 11.5,
    12.0, 12.5, 13.0, 13.5, 14.0, 14.5, 15.0, 15.5, 16.0, 16.5, 17.0, 17.5, 18.0, 18.5, 19.0, 19.5, 20.0, 20.5,
    21.0, 21.5, 22.0, 22.5, 23.0, 23.5, 24.0, 24.5, 25.0, 25.01, 25.02, 50.0};
static const double CP_TAB[NTAB] = {
    0.0, 0.0, 0.0, 0.178085, 0.289075, 0.349022, 0.384728, 0.406059, 0.420228, 0.428823, 0.433873,
    0.436223, 0.436845, 0.436575, 0.436511, 0.436561, 0.436517, 0.435903, 0.434673, 0.433230, 0.430466, 0.378869,
    0.335199, 0.297991, 0.266092, 0.238588, 0.214748, 0.193981, 0.175808, 0.159835, 0.145741, 0.133256, 0.122157,
    0.112257, 0.103399, 0.095449, 0.088294, 0.081836, 0.075993, 0.070692, 0.065875, 0.061484, 0.057476, 0.053809,
    0.050447, 0.047358, 0.044518, 0.041900, 0.039483, 0.0, 0.0};
static const double CT_TAB[NTAB] = {
    0.0, 0.0, 0.0, 0.99, 0.99, 0.97373036, 0.92826162, 0.89210543, 0.86100905, 0.835423, 0.81237673,
    0.79225789, 0.77584769, 0.7629228, 0.76156073, 0.76261984, 0.76169723, 0.75232027, 0.74026851, 0.72987175,
    0.70701647, 0.54054532, 0.45509459, 0.39343381, 0.34250785, 0.30487242, 0.27164979, 0.24361964, 0.21973831,
    0.19918151, 0.18131868, 0.16537679, 0.15103727, 0.13998636, 0.1289037, 0.11970413, 0.11087113, 0.10339901,
    0.09617888, 0.09009926, 0.08395078, 0.0791188, 0.07448356, 0.07050731, 0.06684119, 0.06345518, 0.06032267,
    0.05741999, 0.05472609, 0.0, 0.0};

/* case.yaml / nrel_5MW constants */
static const double D = 126.0, HH = 90.0, TSR = 8.0, PP = 1.88;
static const double SHEAR = 0.12, RHO = 1.225, REF_RHO = 1.225;
static const double ALPHA = 0.58, BETA = 0.077, KA = 0.38, KB = 0.004, AD = 0.0, BD = 0.0, DM = 1.0;
static const double CH_CONST = 0.5, CH_AI = 0.8, CH_INIT = 0.1, CH_DOWN = -0.32;
static const double NUM_EPS = 0.001;
#define PI 3.141592653589793

static double radians(double a) { return a * (PI / 180.0); }
static double degrees(double a) { return a * (180.0 / PI); }
static double cosd(double a) { return cos(radians(a)); }
static double sind(double a) { return sin(radians(a)); }

/* numpy's add.reduce over the 9 contiguous grid values: pairwise over the first 8, then the 9th */
static double sum9(const double* p) {
    return (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) + p[8];
}
static double mean9(const double* p) { return sum9(p) / 9.0; }
static double cubic_mean9(const double* u) {
    double c[NP];
    for (int p = 0; p < NP; ++p) c[p] = u[p] * u[p] * u[p]; /* numpy evaluates u**3 with pow(); differs by <=1 ulp */
    return cbrt(mean9(c));
}
static double std9(const double* a) {
    double m = mean9(a), d[NP];
    for (int p = 0; p < NP; ++p) { double t = a[p] - m; d[p] = t * t; }
    return sqrt(mean9(d));
}

/* np.interp on the 51-row table + scipy fill values outside [xp[0], xp[-1]] */
static double interp_tab(double x, const double* fp, double left, double right) {
    if (x < WS_TAB[0]) return left;
    if (x > WS_TAB[NTAB - 1]) return right;
    if (x == WS_TAB[NTAB - 1]) return fp[NTAB - 1];
    int j = 0;
    while (j + 1 < NTAB - 1 && WS_TAB[j + 1] <= x) ++j; /* xp[j] <= x < xp[j+1] */
    if (x == WS_TAB[j]) return fp[j];
    double slope = (fp[j + 1] - fp[j]) / (WS_TAB[j + 1] - WS_TAB[j]);
    return slope * (x - WS_TAB[j]) + fp[j];
}

static double fmod_py(double a, double m) { /* python/numpy % for positive m */
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

double wf_oracle_pairwise_sum(const double* a, int n);

typedef struct {
    int* order;       /* [T]   sorted position -> original turbine index */
    double* power_W;  /* [T]   original order */
    double* ws_local; /* [T] */
    double* wd_local; /* [T] */
    double* ti;       /* [T] */
    double* std_u;    /* [T] */
    double* std_v;    /* [T] */
    double* std_w;    /* [T] */
    double* u;        /* [T*9] original order, may be NULL */
    double* v;
    double* w;
    double* ti_field;
} OracleOut;

/* One solve.  cs: NULL or {cosd(dev), sind(dev)} override.  order_in: NULL (stable sort) or a host-chosen order. */
int wf_oracle_solve(int T, const double* lx, const double* ly, double ws, double wd, const double* yaw_deg,
                    const double* cs, const int* order_in, double ti_ambient, OracleOut* out) {
    const double I0 = ti_ambient;
    const double eps = 0.2 * D;
    const double off[G] = {-31.5, 0.0, 31.5}; /* linspace(-D/4, D/4, 3) */
    double* buf = (double*)malloc(sizeof(double) * (size_t)T * (4 + 4 * NP + 2 * NP + 2));
    if (!buf) return -1;
    double* xr = buf;
    double* yr = xr + T;
    double* xs = yr + T;
    double* ys = xs + T;
    double* u = ys + T;       /* [T][9] sorted */
    double* v = u + T * NP;
    double* w = v + T * NP;
    double* ti = w + T * NP;
    double* wake = ti + T * NP;
    double* defc = wake + T * NP; /* deficit scratch [T][9] */
    double* yaws = defc + T * NP;
    int* ord = out->order;

    /* A.2 geometry */
    double dev = fmod_py(fmod_py(wd - 270.0, 360.0) + 360.0, 360.0);
    double c = cs ? cs[0] : cosd(dev), s = cs ? cs[1] : sind(dev);
    double xmin = lx[0], xmax = lx[0], ymin = ly[0], ymax = ly[0];
    for (int t = 1; t < T; ++t) {
        if (lx[t] < xmin) xmin = lx[t];
        if (lx[t] > xmax) xmax = lx[t];
        if (ly[t] < ymin) ymin = ly[t];
        if (ly[t] > ymax) ymax = ly[t];
    }
    double xc = (xmin + xmax) / 2, yc = (ymin + ymax) / 2;
    for (int t = 0; t < T; ++t) {
        double xo = lx[t] - xc, yo = ly[t] - yc;
        volatile double a1 = xo * c, a2 = yo * s, b1 = xo * s, b2 = yo * c; /* forbid FMA contraction */
        xr[t] = (a1 - a2) + xc;
        yr[t] = (b1 + b2) + yc;
    }
    if (order_in) {
        for (int t = 0; t < T; ++t) ord[t] = order_in[t];
    } else { /* stable insertion sort by xr */
        for (int t = 0; t < T; ++t) ord[t] = t;
        for (int a = 1; a < T; ++a) {
            int k = ord[a], b = a - 1;
            while (b >= 0 && xr[ord[b]] > xr[k]) { ord[b + 1] = ord[b]; --b; }
            ord[b + 1] = k;
        }
    }
    for (int t = 0; t < T; ++t) { xs[t] = xr[ord[t]]; ys[t] = yr[ord[t]]; yaws[t] = yaw_deg[ord[t]]; }

    /* A.3 initial flow (depends on the vertical grid index k only) */
    double Z[G], U0[G], dU0[G], nu[G];
    for (int k = 0; k < G; ++k) {
        Z[k] = HH + off[k];
        U0[k] = ws * pow(Z[k] / HH, SHEAR);
        dU0[k] = ws * (SHEAR * pow(1 / HH, SHEAR) * pow(Z[k], SHEAR - 1));
        double lmda = D / 8, kappa = 0.41;
        double lm = kappa * Z[k] / (1 + kappa * Z[k] / lmda);
        nu[k] = lm * lm * fabs(dU0[k]);
    }
    /* Uinf = np.mean over all T*9 initial velocities (numpy pairwise summation over contiguous values) */
    double Uinf;
    {
        int n = T * NP;
        double* a = (double*)malloc(sizeof(double) * (size_t)n);
        if (!a) { free(buf); return -1; }
        for (int i = 0; i < n; ++i) a[i] = U0[i % G];
        Uinf = wf_oracle_pairwise_sum(a, n) / n;
        free(a);
    }
    for (int t = 0; t < T; ++t)
        for (int p = 0; p < NP; ++p) {
            u[t * NP + p] = U0[p % G];
            v[t * NP + p] = 0.0;
            w[t * NP + p] = 0.0;
            ti[t * NP + p] = I0;
            wake[t * NP + p] = 0.0;
        }
    const double vel_top = pow((HH + D / 2) / HH, SHEAR), vel_bot = pow((HH - D / 2) / HH, SHEAR);

    /* A.4 - A.8 sequential solver */
    for (int i = 0; i < T; ++i) {
        double Xi[NP], Yi[NP];
        for (int p = 0; p < NP; ++p) { Xi[p] = xs[i]; Yi[p] = ys[i] + off[p / G]; }
        const double x_i = mean9(Xi), y_i = mean9(Yi);
        const double yaw_i = yaws[i];
        const double cy = cosd(yaw_i), sy = sind(yaw_i);
        const double avg = cubic_mean9(&u[i * NP]);
        double ct = interp_tab(avg, CT_TAB, 0.0001, 0.9999);
        ct = fmin(fmax(ct, 0.0001), 0.9999);
        ct = ct * cy * 1.0;
        const double a_i = 0.5 / (cy * 1.0) * (1 - sqrt(1 - ct * cy * 1.0));
        double ti_i[NP];
        for (int p = 0; p < NP; ++p) ti_i[p] = ti[i * NP + p];
        const double G_top0 = (PI / 8) * D * vel_top * Uinf * ct;
        const double G_bot0 = (PI / 8) * D * vel_bot * Uinf * ct;
        const double G_wr = 0.25 * 2 * PI * D * (a_i - a_i * a_i) * avg / TSR;

        /* A.5 secondary steering */
        double eff_yaw;
        {
            double vt[NP], vb[NP], vc[NP];
            for (int p = 0; p < NP; ++p) {
                double yL = (Yi[p] - y_i) + NUM_EPS, z = Z[p % G];
                double zT = z - (HH + D / 2) + NUM_EPS, rT = yL * yL + zT * zT;
                vt[p] = (G_top0 * zT) / (2 * PI * rT) * (1 - exp(-rT / (eps * eps)));
                double zB = z - (HH - D / 2) + NUM_EPS, rB = yL * yL + zB * zB;
                vb[p] = ((-1 * G_bot0) * zB) / (2 * PI * rB) * (1 - exp(-rB / (eps * eps)));
                double zC = z - HH + NUM_EPS, rC = yL * yL + zC * zC;
                vc[p] = (G_wr * zC) / (2 * PI * rC) * (1 - exp(-rC / (eps * eps)));
            }
            double val = 2 * (mean9(&v[i * NP]) - mean9(vc)) / (mean9(vt) + mean9(vb));
            if (val < -1.0) val = -1.0;
            if (val > 1.0) val = 1.0;
            eff_yaw = yaw_i + degrees(0.5 * asin(val));
        }

        /* per-source, per-grid-index deflection parameters (TI_i is indexed by the TARGET's grid index p) */
        const double g = -1 * eff_yaw, cg = cosd(g);
        const double gv = -1 * yaw_i, cgv = cosd(gv);
        const double Gt = sy * cy * G_top0, Gb = -1 * sy * cy * G_bot0;

        /* A.7 transverse velocities on every target (needs the source's own contribution first for the TI update) */
        double* Vw = defc; /* reuse scratch: store V in defc, W in a second scratch */
        double* Ww = (double*)malloc(sizeof(double) * (size_t)T * NP);
        for (int t = 0; t < T; ++t) {
            const double dx = xs[t] - x_i;
            for (int p = 0; p < NP; ++p) {
                const int k = p % G;
                double yL = ((ys[t] + off[p / G]) - y_i) + NUM_EPS, z = Z[k];
                double decay = eps * eps / (4 * nu[k] * dx / Uinf + eps * eps);
                double zz, r, core, V1, W1, V2, W2, V3, W3, V4, W4, V5, W5, V6, W6;
                zz = z - (HH + D / 2) + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V1 = (Gt * zz) / (2 * PI * r) * core * decay; W1 = (-1 * Gt * yL) / (2 * PI * r) * core * decay;
                zz = z - (HH - D / 2) + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V2 = (Gb * zz) / (2 * PI * r) * core * decay; W2 = (-1 * Gb * yL) / (2 * PI * r) * core * decay;
                zz = z - HH + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V5 = (G_wr * zz) / (2 * PI * r) * core * decay; W5 = (-1 * G_wr * yL) / (2 * PI * r) * core * decay;
                zz = z + (HH + D / 2) + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V3 = (-1 * Gt * zz) / (2 * PI * r) * core * decay; W3 = (Gt * yL) / (2 * PI * r) * core * decay;
                zz = z + (HH - D / 2) + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V4 = (-1 * Gb * zz) / (2 * PI * r) * core * decay; W4 = (Gb * yL) / (2 * PI * r) * core * decay;
                zz = z + HH + NUM_EPS; r = yL * yL + zz * zz; core = 1 - exp(-r / (eps * eps));
                V6 = (-1 * G_wr * zz) / (2 * PI * r) * core * decay; W6 = (G_wr * yL) / (2 * PI * r) * core * decay;
                double V = ((((V1 + V2) + V3) + V4) + V5) + V6;
                double W = ((((W1 + W2) + W3) + W4) + W5) + W6;
                if (dx < 0.0) { V = 0.0; W = 0.0; }
                if (W < 0.0) W = 0.0;
                Vw[t * NP + p] = V;
                Ww[t * NP + p] = W;
            }
        }

        /* A.6 deflection needs TI_i BEFORE the yaw-added-recovery update; A.8 needs it AFTER */
        double x0d[NP], kyd[NP], delta0[NP], farK[NP], sM0[NP], sy0d[NP], sz0d[NP];
        for (int p = 0; p < NP; ++p) {
            const double U = U0[p % G];
            double uR = U * ct * cg / (2.0 * (1 - sqrt(1 - (ct * cg))));
            double u0 = U * sqrt(1 - ct);
            x0d[p] = D * (cg * (1 + sqrt(1 - ct * cg))) / (sqrt(2.0) * (4 * ALPHA * ti_i[p] + 2 * BETA * (1 - sqrt(1 - ct)))) + x_i;
            kyd[p] = KA * ti_i[p] + KB;
            double C0 = 1 - u0 / U, M0 = C0 * (2 - C0);
            double E0 = C0 * C0 - 3 * exp(1.0 / 12.0) * C0 + 3 * exp(1.0 / 3.0);
            sz0d[p] = D * 0.5 * sqrt(uR / (U + u0));
            sy0d[p] = sz0d[p] * cg * cosd(0.0);
            double th = DM * (0.3 * radians(g) / cg);
            th = th * (1 - sqrt(1 - ct * cg));
            delta0[p] = tan(th) * (x0d[p] - x_i);
            sM0[p] = sqrt(M0);
            farK[p] = th * E0 / 5.2 * sqrt(sy0d[p] * sz0d[p] / (kyd[p] * kyd[p] * M0));
        }

        /* yaw-added recovery: in-place TI update of the source */
        {
            double I = ti_i[0];
            double k = (avg * I) * (avg * I) / (2.0 / 3.0);
            double u_term = sqrt(2 * k);
            double tv[NP], tw[NP];
            for (int p = 0; p < NP; ++p) { tv[p] = v[i * NP + p] + Vw[i * NP + p]; tw[p] = w[i * NP + p] + Ww[i * NP + p]; }
            double v_term = mean9(tv), w_term = mean9(tw);
            double k_total = 0.5 * (u_term * u_term + v_term * v_term + w_term * w_term);
            double I_total = sqrt((2.0 / 3.0) * k_total) / avg;
            double I_mix = I_total - I;
            for (int p = 0; p < NP; ++p) { ti_i[p] = ti_i[p] + 2 * I_mix; ti[i * NP + p] = ti_i[p]; }
        }

        /* A.8 per-source deficit parameters with the UPDATED TI */
        double x0v[NP], kyv[NP], sy0v[NP], sz0v[NP];
        for (int p = 0; p < NP; ++p) {
            const double U = U0[p % G];
            double uR = U * ct / (2.0 * (1 - sqrt(1 - ct)));
            double u0 = U * sqrt(1 - ct);
            sz0v[p] = D * 0.5 * sqrt(uR / (U + u0));
            sy0v[p] = sz0v[p] * cgv * cosd(0.0);
            double x0 = 1.0 * (D * cgv * (1 + sqrt(1 - ct)));
            x0 = x0 / (sqrt(2.0) * (4 * ALPHA * ti_i[p] + 2 * BETA * (1 - sqrt(1 - ct))));
            x0v[p] = x0 + x_i;
            kyv[p] = KA * ti_i[p] + KB;
        }

        for (int t = 0; t < T; ++t) {
            const double X = xs[t];
            double dU[NP];
            int cnt = 0;
            for (int p = 0; p < NP; ++p) {
                const int k = p % G;
                const double Y = ys[t] + off[p / G], z = Z[k], U = U0[k];
                /* deflection */
                double dn = ((X - x_i) / (x0d[p] - x_i)) * delta0[p] + (AD + BD * (X - x_i));
                dn = dn * (X >= x_i ? 1.0 : 0.0);
                dn = dn * (X <= x0d[p] ? 1.0 : 0.0);
                double sgy = kyd[p] * (X - x0d[p]) + sy0d[p], sgz = kyd[p] * (X - x0d[p]) + sz0d[p];
                sgy = sgy * (X >= x0d[p] ? 1.0 : 0.0) + sy0d[p] * (X < x0d[p] ? 1.0 : 0.0);
                sgz = sgz * (X >= x0d[p] ? 1.0 : 0.0) + sz0d[p] * (X < x0d[p] ? 1.0 : 0.0);
                double sq = sqrt(sgy * sgz / (sy0d[p] * sz0d[p]));
                double lnn = (1.6 + sM0[p]) * (1.6 * sq - sM0[p]);
                double lnd = (1.6 - sM0[p]) * (1.6 * sq + sM0[p]);
                double df = delta0[p] + farK[p] * log(lnn / lnd) + (AD + BD * (X - x_i));
                df = df * (X > x0d[p] ? 1.0 : 0.0);
                double defl = dn + df;
                /* deficit */
                double deficit = 0.0;
                const double x0 = x0v[p];
                const int near = (X > x_i + 0.1) && (X < x0);
                const int far = (X >= x0);
                double dy = Y - y_i - defl, dz = z - HH;
                if (near) {
                    double up = (X - x_i) / (x0 - x_i), down = (x0 - X) / (x0 - x_i);
                    double s_y = down * 0.501 * D * sqrt(ct / 2.0) + up * sy0v[p];
                    double s_z = down * 0.501 * D * sqrt(ct / 2.0) + up * sz0v[p];
                    double a = 1.0 / (2 * s_y * s_y) + 0.0, cc = 0.0 + 1.0 / (2 * s_z * s_z);
                    double r = a * (dy * dy) - 0.0 + cc * (dz * dz);
                    double d = 1 - (ct * cgv / (8.0 * s_y * s_z / (D * D)));
                    d = fmin(fmax(d, 0.0), 1.0);
                    deficit += (1 - sqrt(d)) * exp(-1 * r / (2 * sqrt(0.5) * sqrt(0.5)));
                }
                if (far) {
                    double s_y = (kyv[p] * (X - x0) + sy0v[p]), s_z = (kyv[p] * (X - x0) + sz0v[p]);
                    double a = 1.0 / (2 * s_y * s_y) + 0.0, cc = 0.0 + 1.0 / (2 * s_z * s_z);
                    double r = a * (dy * dy) - 0.0 + cc * (dz * dz);
                    double d = 1 - (ct * cgv / (8.0 * s_y * s_z / (D * D)));
                    d = fmin(fmax(d, 0.0), 1.0);
                    deficit += (1 - sqrt(d)) * exp(-1 * r / (2 * sqrt(0.5) * sqrt(0.5)));
                }
                dU[p] = deficit * U;
                if (dU[p] > 0.05) ++cnt;
            }
            /* SOSFS, Crespo-Hernandez, field updates */
            const double dx = X - x_i;
            const double upm = dx <= 0.1 ? 1.0 : 0.0, dnm = dx > -0.1 ? 1.0 : 0.0;
            const double dxp = dx * dnm + 1.0 * upm;
            double wat = CH_CONST * pow(a_i, CH_AI) * pow(I0, CH_INIT) * pow(dxp / D, CH_DOWN);
            wat = wat * dnm;
            if (isinf(wat) && wat > 0) wat = 0.0;
            if (isnan(wat)) wat = 0.0;
            const double overlap = (double)cnt / (G * G);
            for (int p = 0; p < NP; ++p) {
                const double Y = ys[t] + off[p / G];
                double ti_add = overlap * wat * (X > x_i ? 1.0 : 0.0) * (fabs(y_i - Y) < 2 * D ? 1.0 : 0.0)
                                * (X <= 15 * D + x_i ? 1.0 : 0.0);
                double cand = sqrt(ti_add * ti_add + I0 * I0);
                int idx = t * NP + p;
                ti[idx] = fmax(cand, ti[idx]);
                wake[idx] = hypot(wake[idx], dU[p]);
                u[idx] = U0[p % G] - wake[idx];
                v[idx] = v[idx] + Vw[idx];
                w[idx] = w[idx] + Ww[idx];
            }
        }
        free(Ww);
    }

    /* A.9 / A.10 finalise in ORIGINAL order */
    for (int t = 0; t < T; ++t) {
        const int o = ord[t];
        const double* ut = &u[t * NP];
        const double* vt = &v[t * NP];
        const double* wt = &w[t * NP];
        double avg = cubic_mean9(ut);
        double veff = pow(RHO / REF_RHO, 1.0 / 3.0) * avg * pow(cosd(yaws[t]), PP / 3.0) * pow(cosd(0.0), PP / 3.0);
        /* inner power table: 0.5 * pi * R^2 * Cp * eta * ws^3, built like numpy does */
        double ptab[NTAB];
        const double area = PI * pow(D / 2.0, 2.0);
        for (int k = 0; k < NTAB; ++k) ptab[k] = 0.5 * area * CP_TAB[k] * 1.0 * (WS_TAB[k] * WS_TAB[k] * WS_TAB[k]);
        out->power_W[o] = interp_tab(veff, ptab, 0.0, 0.0) * REF_RHO;
        out->ws_local[o] = avg;
        double dd[NP];
        for (int p = 0; p < NP; ++p) dd[p] = wd - degrees(atan2(vt[p], ut[p]));
        out->wd_local[o] = mean9(dd);
        out->ti[o] = mean9(&ti[t * NP]);
        out->std_u[o] = std9(ut);
        out->std_v[o] = std9(vt);
        out->std_w[o] = std9(wt);
        if (out->u) memcpy(&out->u[o * NP], ut, sizeof(double) * NP);
        if (out->v) memcpy(&out->v[o * NP], vt, sizeof(double) * NP);
        if (out->w) memcpy(&out->w[o * NP], wt, sizeof(double) * NP);
        if (out->ti_field) memcpy(&out->ti_field[o * NP], &ti[t * NP], sizeof(double) * NP);
    }
    free(buf);
    return 0;
}

/* numpy's pairwise summation (npy_math pairwise_sum for contiguous doubles) */
double wf_oracle_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return wf_oracle_pairwise_sum(a, n2) + wf_oracle_pairwise_sum(a + n2, n - n2);
    }
}

/* Batched driver (pthreads): B independent envs sharing one layout. Arrays are [B][T] row-major. */
#include <pthread.h>

typedef struct {
    int B, T, nthreads, tid;
    const double *lx, *ly, *ws, *wd, *yaw, *cs;
    double ti_ambient;
    int* order;
    double *power_W, *ws_local, *wd_local, *ti, *std_u, *std_v, *std_w;
    int rc;
} BatchJob;

static void* batch_worker(void* arg) {
    BatchJob* j = (BatchJob*)arg;
    for (int b = j->tid; b < j->B; b += j->nthreads) {
        OracleOut o;
        size_t k = (size_t)b * j->T;
        o.order = j->order + k; o.power_W = j->power_W + k; o.ws_local = j->ws_local + k;
        o.wd_local = j->wd_local + k; o.ti = j->ti + k; o.std_u = j->std_u + k; o.std_v = j->std_v + k;
        o.std_w = j->std_w + k;
        o.u = o.v = o.w = o.ti_field = NULL;
        int r = wf_oracle_solve(j->T, j->lx, j->ly, j->ws[b], j->wd[b], j->yaw + k, j->cs ? j->cs + 2 * b : NULL,
                                NULL, j->ti_ambient, &o);
        if (r) j->rc = r;
    }
    return NULL;
}

int wf_oracle_solve_batch(int B, int T, const double* lx, const double* ly, const double* ws, const double* wd,
                          const double* yaw_deg, const double* cs /* [B][2] or NULL */, double ti_ambient,
                          int nthreads, int* order, double* power_W, double* ws_local, double* wd_local,
                          double* ti, double* std_u, double* std_v, double* std_w) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    BatchJob jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        BatchJob j = {B, T, nthreads, t, lx, ly, ws, wd, yaw_deg, cs, ti_ambient, order,
                      power_W, ws_local, wd_local, ti, std_u, std_v, std_w, 0};
        jobs[t] = j;
        if (nthreads == 1) batch_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    return rc;
}
