// Host-only helpers shared by the C-ABI layer and the build-time constant generator (gen_baked.cu):
// the reference defaults (case.yaml + nrel_5MW) and the FP64 precomputation of the tuned kernel's constants.
#pragma once
#include "../../include/wfcrl_b200.h"
#include "wf_device.cuh"

#include <math.h>
#include <string.h>

static inline void wf_fill_default_config(WfConfig* c) {
    memset(c, 0, sizeof(*c));
    c->precision = WF_PREC_F64;
    c->kernel = WF_KERNEL_BASIC;
    c->continuous_control = 1;
    c->reward_shaper = WF_SHAPER_NONE;
    c->yaw_lo = -40.0; c->yaw_hi = 40.0; c->yaw_step = 5.0;  // data_cases.py:21
    c->load_coef = 0.1;                                       // simple_env.py:25
    c->shaper_reference = 0.0;
    c->dt = 60.0;                                             // data_cases.py:507
    c->actuator_rate = 0.3;                                   // mdp.py:52
    c->air_density = 1.225; c->turbulence_intensity = 0.06; c->wind_shear = 0.12; c->wind_veer = 0.0;
    c->alpha = 0.58; c->beta = 0.077; c->ka = 0.38; c->kb = 0.004; c->ad = 0.0; c->bd = 0.0; c->dm = 1.0;
    c->ch_initial = 0.1; c->ch_constant = 0.5; c->ch_ai = 0.8; c->ch_downstream = -0.32;
    c->turbine_grid_points = 3;                               // case.yaml:16
    c->rotor_diameter = 126.0; c->hub_height = 90.0; c->tsr = 8.0; c->pP = 1.88; c->pT = 1.88;
    c->generator_efficiency = 1.0; c->ref_density_cp_ct = 1.225;
    // nrel_5MW power/thrust table as shipped with FLORIS 3.x
    static const double cp[51] = {
        0.0, 0.0, 0.0, 0.178085, 0.289075, 0.349022, 0.384728, 0.406059, 0.420228, 0.428823, 0.433873, 0.436223,
        0.436845, 0.436575, 0.436511, 0.436561, 0.436517, 0.435903, 0.434673, 0.433230, 0.430466, 0.378869, 0.335199,
        0.297991, 0.266092, 0.238588, 0.214748, 0.193981, 0.175808, 0.159835, 0.145741, 0.133256, 0.122157, 0.112257,
        0.103399, 0.095449, 0.088294, 0.081836, 0.075993, 0.070692, 0.065875, 0.061484, 0.057476, 0.053809, 0.050447,
        0.047358, 0.044518, 0.041900, 0.039483, 0.0, 0.0};
    static const double ct[51] = {
        0.0, 0.0, 0.0, 0.99, 0.99, 0.97373036, 0.92826162, 0.89210543, 0.86100905, 0.835423, 0.81237673, 0.79225789,
        0.77584769, 0.7629228, 0.76156073, 0.76261984, 0.76169723, 0.75232027, 0.74026851, 0.72987175, 0.70701647,
        0.54054532, 0.45509459, 0.39343381, 0.34250785, 0.30487242, 0.27164979, 0.24361964, 0.21973831, 0.19918151,
        0.18131868, 0.16537679, 0.15103727, 0.13998636, 0.1289037, 0.11970413, 0.11087113, 0.10339901, 0.09617888,
        0.09009926, 0.08395078, 0.0791188, 0.07448356, 0.07050731, 0.06684119, 0.06345518, 0.06032267, 0.05741999,
        0.05472609, 0.0, 0.0};
    c->table_len = 51;
    c->table_ws[0] = 0.0; c->table_ws[1] = 2.0; c->table_ws[2] = 2.5;
    for (int i = 0; i < 45; ++i) c->table_ws[3 + i] = 3.0 + 0.5 * i;
    c->table_ws[48] = 25.01; c->table_ws[49] = 25.02; c->table_ws[50] = 50.0;
    for (int i = 0; i < 51; ++i) { c->table_cp[i] = cp[i]; c->table_ct[i] = ct[i]; }
    }

// Host-side (FP64) precomputation of the tuned FP32 kernel's per-model constants.
template <typename R> static inline void build_fast_const(const WfConfig& c, WfFastConstT<R>* f) {
    memset(f, 0, sizeof(*f));
    const double PI = 3.141592653589793, NUM_EPS = 0.001;
    const double D = c.rotor_diameter, HH = c.hub_height, eps = 0.2 * D, eps2 = eps * eps, off = 0.5 * D / 2;
    double Z[3], ratio[3], rsum = 0.0;
    for (int k = 0; k < 3; ++k) {
        Z[k] = HH + (k - 1) * off;
        ratio[k] = pow(Z[k] / HH, c.wind_shear);
        rsum += ratio[k];
    }
    const double mean_ratio = rsum / 3.0;
    f->mean_ratio = (R)mean_ratio;
    const double zc[6] = {-(HH + D / 2), -(HH - D / 2), (HH + D / 2), (HH - D / 2), -HH, HH};
    double zz[6][3];
    for (int k = 0; k < 3; ++k) {
        f->ratio[k] = (R)ratio[k];
        const double dU = c.wind_shear * pow(1.0 / HH, c.wind_shear) * pow(Z[k], c.wind_shear - 1.0);  // per unit ws
        const double lmda = D / 8, kappa = 0.41;
        const double lm = kappa * Z[k] / (1 + kappa * Z[k] / lmda);
        f->nu4[k] = (R)(4 * lm * lm * fabs(dU) / mean_ratio);
        for (int q = 0; q < 6; ++q) {
            zz[q][k] = (Z[k] + zc[q]) + NUM_EPS;
            f->zz[q][k] = (R)zz[q][k];
            f->zz2[q][k] = (R)(zz[q][k] * zz[q][k]);
            f->ez[q][k] = (R)exp(-zz[q][k] * zz[q][k] / eps2);
        }
        f->dz2[k] = (R)(((k - 1) * off) * ((k - 1) * off));
        f->offj[k] = (R)((k - 1) * off);
    }
    // packed block: per k four float4 = (zz2_top, zz2_topmirror, zz2_bot, zz2_botmirror) (zz2_core, zz2_coremirror,
    // ez_top, ez_bot) (ez_core, zz_top, zz_topmirror, zz_bot) (zz_botmirror, zz_core, zz_coremirror, nu4).
    // Vortex order in zz[][]: 0 top, 1 bottom, 2 top mirror, 3 bottom mirror, 4 core, 5 core mirror.  The mirror vortices'
    // cores are taken as 1 (exp(-zz^2/eps^2) <= 1.1e-5 for them).
    for (int k = 0; k < 3; ++k) {
        R* c = f->cblk + 16 * k;
        c[0] = f->zz2[0][k]; c[1] = f->zz2[2][k]; c[2] = f->zz2[1][k]; c[3] = f->zz2[3][k];
        c[4] = f->zz2[4][k]; c[5] = f->zz2[5][k]; c[6] = f->ez[0][k]; c[7] = f->ez[1][k];
        c[8] = f->ez[4][k]; c[9] = f->zz[0][k]; c[10] = f->zz[2][k]; c[11] = f->zz[1][k];
        c[12] = f->zz[3][k]; c[13] = f->zz[4][k]; c[14] = f->zz[5][k]; c[15] = f->nu4[k];
    }
    double a_top = 0, a_bot = 0, a_core = 0, sv[3] = {0, 0, 0};
    for (int p = 0; p < 9; ++p) {
        const int j = p / 3, k = p % 3;
        const double yL = (j - 1) * off + NUM_EPS, q = yL * yL;
        double fq[6];
        for (int v = 0; v < 6; ++v) {
            const double r = q + zz[v][k] * zz[v][k];
            fq[v] = (1 - exp(-r / eps2)) / (2 * PI * r);
        }
        a_top += zz[0][k] * fq[0] / 9.0;
        a_bot += zz[1][k] * fq[1] / 9.0;
        a_core += zz[4][k] * fq[4] / 9.0;
        const double cv[3] = {zz[0][k] * fq[0] - zz[2][k] * fq[2], zz[1][k] * fq[1] - zz[3][k] * fq[3],
                              zz[4][k] * fq[4] - zz[5][k] * fq[5]};
        const double cw[3] = {-yL * (fq[0] - fq[2]), -yL * (fq[1] - fq[3]), -yL * (fq[4] - fq[5])};
        for (int v = 0; v < 3; ++v) {
            f->cv[v][p] = (R)cv[v];
            f->cw[v][p] = (R)cw[v];
            sv[v] += cv[v];
        }
    }
    f->a_top = (R)a_top; f->a_bot = (R)a_bot; f->a_core = (R)a_core;
    for (int v = 0; v < 3; ++v) f->sv[v] = (R)sv[v];
    f->D = (R)D; f->inv_D = (R)(1.0 / D); f->eps2 = (R)eps2; f->inv_eps2 = (R)(1.0 / eps2);
    f->inv_2pi = (R)(1.0 / (2 * PI));
    const double vel_top = pow((HH + D / 2) / HH, c.wind_shear), vel_bot = pow((HH - D / 2) / HH, c.wind_shear);
    f->c_top = (R)((PI / 8) * D * vel_top * mean_ratio);
    f->c_bot = (R)((PI / 8) * D * vel_bot * mean_ratio);
    f->c_wr = (R)(0.25 * 2 * PI * D / c.tsr);
    f->inv_ss_den = (R)(1.0 / (((PI / 8) * D * vel_top * mean_ratio) * a_top - ((PI / 8) * D * vel_bot * mean_ratio) * a_bot));
    f->alpha4 = (R)(4 * c.alpha); f->beta2 = (R)(2 * c.beta); f->ka = (R)c.ka; f->kb = (R)c.kb;
    f->ad = (R)c.ad; f->bd = (R)c.bd; f->dm03 = (R)(0.3 * c.dm);
    f->e3_112 = (R)(3 * exp(1.0 / 12.0)); f->e3_13 = (R)(3 * exp(1.0 / 3.0));
    f->near_c = (R)(0.501 * D * sqrt(0.5));
    f->d2_8 = (R)(D * D / 8.0);
    f->ch_const = (R)c.ch_constant; f->ch_ai = (R)c.ch_ai; f->ch_init = (R)c.ch_initial;
    f->ch_down = (R)c.ch_downstream;
    f->pP3 = (R)(c.pP / 3.0); f->rho_fac = (R)cbrt(c.air_density / c.ref_density_cp_ct);
    f->ref_rho = (R)c.ref_density_cp_ct; f->two_D = (R)(2 * D);
    f->load_coef = (R)c.load_coef; f->shaper_reference = (R)c.shaper_reference;
    const int n = c.table_len;
    f->table_len = n;
    const double area = PI * pow(D / 2.0, 2.0);
    for (int i = 0; i < n; ++i) {
        f->tab_ws[i] = (R)c.table_ws[i];
        f->tab_ct[i] = (R)c.table_ct[i];
        f->tab_pw[i] = (R)(0.5 * area * c.table_cp[i] * c.generator_efficiency * pow(c.table_ws[i], 3.0));
    }
    f->coarse_len = 128;
    const double span = c.table_ws[n - 1] - c.table_ws[0];
    f->coarse_scale = (R)(128.0 / span);
    for (int bkt = 0; bkt < 128; ++bkt) {
        // conservative (slightly early) left edge so that float rounding of the bucket index can never skip a node
        const double left = c.table_ws[0] + (bkt - 0.01) * span / 128.0;
        int idx = 0;
        while (idx + 1 < n - 1 && c.table_ws[idx + 1] <= left) ++idx;
        f->coarse[bkt] = (unsigned char)idx;
    }
}

