// Internal device-side structures shared by the kernels and the C-ABI layer (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WF_MAX_TURBINES_K 128  // == WF_MAX_TURBINES of the public header
#define WF_NP 9  // 3x3 rotor grid (case.yaml:16), p = 3*j + k with j lateral, k vertical

// Model constants, passed to kernels by value (kernel parameter space = constant bank, broadcast reads).
struct WfModel {
    int T, B;
    int max_iter, continuous, multi_agent, shaper, table_len;
    float yaw_lo_f, yaw_hi_f, yaw_step_f;  // float32 bounds exactly as gymnasium Box stores them (mdp.py:111-116,143-144)
    float rate_f, dt_f;
    double load_coef, shaper_reference;
    double rho, ref_rho, shear, D, HH, TSR, pP;
    double alpha, beta, ka, kb, ad, bd, dm;
    double ch_const, ch_ai, ch_init, ch_down;
    double e3_112, e3_13;          // 3*exp(1/12), 3*exp(1/3) (Gauss deflection E0)
    double xc, yc;                 // centre of the layout bounding box (rotation centre)
    const double* tab_ws;          // [table_len] device
    const double* tab_ct;          // [table_len]
    const double* tab_pw;          // [table_len] 0.5*area*Cp*eta*ws^3 (power / density)
    const double* layout_x;        // [T] device
    const double* layout_y;        // [T]
};

// Per-env state, device pointers (owned by the handle).
struct WfState {
    double* yaw;        // [B][T] current yaw command / state, degrees (float32-representable in env mode)
    float* acc;         // [B][T] actuation accumulator (mdp.py:157-160, 317-318)
    float* acc_prev;    // [B][T] accumulator one joint action earlier (multi-agent staleness)
    int* num_iter;      // [B] FlorisInterface._num_iter
    int* num_moves;     // [B] WindFarmEnv.num_moves
    double* ws;         // [B] free-stream wind speed
    double* wd;         // [B] free-stream wind direction (already % 360)
    double* ws_norm;    // [B] free-stream speed of the PREVIOUS state (reward normalisation, simple_env.py:79)
    double* shaper_ref; // [B] StepPercentage.reference
    double* ti_amb;     // [B] ambient turbulence intensity
    // geometry (rebuilt when the wind direction changes)
    double* xs;         // [B][T] rotated x, sorted ascending (stable)
    double* ys;         // [B][T] rotated y in the same order
    double* xi;         // [B][T] np.mean of the 9 identical grid x values = fl(fl(8x+x)/9)
    double* yi;         // [B][T] np.mean of the 9 grid y values (numpy reduction order)
    int* order;         // [B][T] sorted position -> original turbine index
    double* cs;         // [B][2] cosd/sind of the deviation from west actually used
};

struct WfOutPtrs {
    void* yaw;
    void* wind_speed;
    void* wind_direction;
    void* power;
    void* load;
    void* reward;
    void* freewind;
    uint8_t* truncated;
};

enum WfMode { WF_MODE_INTERFACE = 0, WF_MODE_ENV = 1, WF_MODE_WARMUP = 2 };

// launchers implemented in wf_kernels.cu
cudaError_t wf_launch_geometry(const WfModel& m, const WfState& s, const uint8_t* d_mask, const double* d_cs_override,
                               cudaStream_t stream);
cudaError_t wf_launch_step_basic(int precision, int mode, const WfModel& m, const WfState& s, const uint8_t* d_mask,
                                 const float* d_action, const double* d_yaw_cmd, const WfOutPtrs& out,
                                 cudaStream_t stream);
cudaError_t wf_launch_reset_state(const WfModel& m, const WfState& s, const uint8_t* d_mask, const double* d_ws,
                                  const double* d_wd, cudaStream_t stream);
cudaError_t wf_step_basic_attributes(int precision, cudaFuncAttributes* attr, int* ctas_per_sm, int threads);
