/*
 * Plain-C client of libwfcrl_b200.so: no Python, no torch, host buffers only.
 *
 * Reproduces the reference notebook's reset observation for Ablaincourt_Floris (examples/demo.ipynb:137-138: wind
 * 6.48958384 m/s from 266.363907 deg, zero yaw) through the interface-mode entry points, then takes one env step with a
 * -5 degree command on turbine 1 through the env-mode entry point.
 *
 * Build (from the repo root):
 *   gcc -O2 -Iinclude examples/c_abi_example.c -Lwfcrl_b200 -lwfcrl_b200 -Wl,-rpath,$PWD/wfcrl_b200 -lm -o /tmp/c_abi_example
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "wfcrl_b200.h"

#define T 7
#define CHECK(call)                                                          \
    do {                                                                     \
        if ((call) != WF_OK) {                                               \
            fprintf(stderr, "%s failed: %s\n", #call, wf_last_error());      \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    /* wfcrl/environments/data_cases.py: the Ablaincourt layout */
    const double x[T] = {484.8, 797.1, 1038.8, 1377.6, 1716.9, 2057.3, 2400.0};
    const double y[T] = {274.0, 251.0, 66.9, -22.7, -112.5, -195.3, -259.0};
    WfConfig cfg;
    CHECK(wf_default_config(&cfg));
    cfg.num_turbines = T;
    cfg.num_envs = 1;
    cfg.max_iter = 70;
    cfg.precision = WF_PREC_F64;
    cfg.kernel = WF_KERNEL_FAST;
    WfHandle h = NULL;
    CHECK(wf_create(&cfg, x, y, &h));

    double ws = 6.48958384, wd = 266.363907;
    double dev = fmod(fmod(wd - 270.0, 360.0) + 360.0, 360.0) * (3.14159265358979323846 / 180.0);
    double c = cos(dev), s = sin(dev);
    /* FlorisInterface.init + the warm-up update_command() of WindFarmMDP.reset */
    CHECK(wf_reset(h, NULL, 1, &ws, &wd, &c, &s, 1, NULL, NULL));

    double yaw[T], wsl[T], wdl[T], power[T], load[T][4], reward = 0.0, free_wind[2];
    unsigned char truncated = 0;
    WfHostOut out = {yaw, wsl, wdl, power, load, NULL, free_wind, &truncated};
    CHECK(wf_update_command_host(h, NULL, &out)); /* update_command() with no argument */
    printf("local wind speed:");
    for (int t = 0; t < T; ++t) printf(" %.8f", wsl[t]);
    printf("\nlocal wind direction:");
    for (int t = 0; t < T; ++t) printf(" %.8f", wdl[t]);
    printf("\n");

    float action[T] = {-5.0f, 0, 0, 0, 0, 0, 0};
    out.reward = &reward;
    CHECK(wf_step_host(h, action, &out, NULL, NULL)); /* one WindFarmEnv.step */
    printf("yaw after step:");
    for (int t = 0; t < T; ++t) printf(" %.1f", yaw[t]);
    printf("\npower MW:");
    for (int t = 0; t < T; ++t) printf(" %.9f", power[t]);
    printf("\nreward: %.12f truncated: %d launches: %llu\n", reward, (int)truncated,
           (unsigned long long)wf_launch_count(h));
    CHECK(wf_destroy(h));
    return 0;
}
