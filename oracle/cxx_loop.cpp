// cxx_loop.cpp - TEST / BENCH INFRASTRUCTURE (not part of the product; nothing under mpc-code_b200/ loads it).
//
// "Same algorithm, C++": the host build (g++ -O3 -fopenmp) of the solver code in mpc-code_b200/csrc/*.cuh driving the
// closed loop of MPC_code.py:485-875 for B independent instances, OpenMP over instances.  bench.py times it as the CPU
// arm (`--impl reference`, `cpu_baseline`): the fastest CPU implementation of this path that exists in the repo
// (the reference's own CasADi/IPOPT stack cannot be installed here).  It shares its arithmetic with the device code, so
// it is a BASELINE, never a parity checker - parity is checked against the independent NumPy oracle (oracle/ipm.py,
// oracle/nlp.py), and tests/test_cxx_loop.py checks this loop against that oracle too.
//
// Per step and instance (statement order of MPC_code.py): plant output + noise (:531-541), estimator (:546-668),
// target problem with its cold guess (:690-718), warm start (:740-764), OCP (:769-781), extraction or the
// infeasible fallback (:786-805), plant step (:813-816).  Time-varying parameters px/py are zero (none of the
// benchmarked configurations defines def_px/def_py).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "mpcb_target.cuh"

#if MPCB_HAS_OCP && MPCB_HAS_TARGET && !MPCB_PLANT_NOMINAL

static IpmOpts mk_opts(int max_iter) {
    IpmOpts o; o.max_iter = max_iter; o.tol = 1e-8; o.mu_init = 0.1; o.bound_relax = 1e-8; o.bound_push = 1e-2;
    o.acceptable_tol = 1e-6; o.honor_original_bounds = 0; o.acceptable_iter = 15; return o;
}

// one OCP solve: the tick schedule of mpcb_ocp for a single instance
static void solve_ocp(OcpInst& I, const OcpShared& S, double* scratch, int max_iter) {
    InstState& st = *I.st;
    for (int k = 0; k <= NH; ++k) ocp_init_stage(I, S, k);
    const int max_ticks = (max_iter + 2) * 8;
    for (int tick = 0; tick < max_ticks && st.state != ST_DONE; ++tick) {
        if (st.state == ST_EVAL) { for (int k = 0; k < NH; ++k) ocp_eval_stage(I, S, k); ocp_kkt(I, S, scratch); }
        if (st.state == ST_LS) { for (int k = 0; k < NH; ++k) ocp_trial_stage(I, S, k); ocp_accept(I, S); }
    }
    for (int k = 0; k <= NH; ++k) ocp_export_stage(I, k);
}

extern "C" {

// Launchers (torch.distributed.run) export OMP_NUM_THREADS=1 to their workers: the CPU arm sets its thread count itself.
void cxx_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int cxx_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Closed loop of `nsim` steps for B instances.  x0p [B,NXP] plant states (in/out), x0m [B,NX], u0 [NU], dhat0 [ND],
// P0 [NXI*NXI], noise [nsim,B,NY] or null, sp [nsim, NU+NY+NX] = (usp, ysp, xsp) per step, bounds in the reference
// layouts (ocp_lbx/ubx [NW], ocp_lbg/ubg [NH*NG] stage-major, ss_lbx/ubx [NWS]), estimator constants.
// Outputs: U [nsim,B,NU], iters/status of the OCP [nsim,B]; timed_s [B] (may be null) = seconds each instance spent in
// its steps nwarm..nsim-1 (the steps after its cold start).  Returns the number of OpenMP threads used.
int cxx_closed_loop(int B, int nsim, double h_step, double* xp, const double* x0m, const double* u0, const double* dhat0,
                    const double* P0, const double* noise, const double* sp,
                    const double* ocp_lbx, const double* ocp_ubx, const double* ocp_lbg, const double* ocp_ubg,
                    const double* ss_lbx, const double* ss_ubx, int est_type, const double* Qkf, const double* Rkf,
                    const double* Kest, const double* dmin, const double* dmax, int has_dbounds, int max_iter_ss,
                    int max_iter_dyn, double* U_out, int* iters_out, int* status_out, int nwarm, double* timed_s) {
    std::vector<double> lbi(NWI, -INFINITY), ubi(NWI, INFINITY);
    for (int k = 0; k <= NH; ++k) {
        for (int i = 0; i < NX; ++i) { lbi[k * NZA + i] = ocp_lbx[k * NZ + i]; ubi[k * NZA + i] = ocp_ubx[k * NZ + i]; }
        if (k < NH) for (int i = 0; i < NU; ++i) { lbi[k * NZA + NXA + i] = ocp_lbx[k * NZ + NX + i]; ubi[k * NZA + NXA + i] = ocp_ubx[k * NZ + NX + i]; }
    }
    OcpShared S; S.lbx = lbi.data(); S.ubx = ubi.data(); S.lbg = ocp_lbg; S.ubg = ocp_ubg; S.o = mk_opts(max_iter_dyn);
    TgtShared T; T.lbx = ss_lbx; T.ubx = ss_ubx; T.o = mk_opts(max_iter_ss);
    EstShared E; E.Q = Qkf; E.R = Rkf; E.K = Kest; E.dmin = dmin; E.dmax = dmax; E.has_dbounds = has_dbounds;
    int nthreads = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
#pragma omp single
        nthreads = omp_get_num_threads();
#endif
        std::vector<double> ws(OcpLayout::total, 0.0), w(NW), wguess(NW), wopt(NW), par(NPAR, 0.0), parss(MPCB_NPARSS, 0.0);
        double scratch[KktScratch::total];
        InstState st;
#pragma omp for schedule(dynamic, 1)
        for (int inst = 0; inst < B; ++inst) {
            double xi[NXI], P[NXI * NXI], u[NU], us[NU], xs[NX], x0[NX], zero_px[NPX + 1] = {0}, zero_py[NPY + 1] = {0};
            double zp[MPCB_NPXP + MPCB_NPYP + 1] = {0};
            double* x = xp + (size_t)inst * MPCB_NXP;
            for (int i = 0; i < NX; ++i) { x0[i] = x0m[(size_t)inst * NX + i]; xi[i] = x0[i]; xs[i] = x0[i]; }
            for (int i = NX; i < NXI; ++i) xi[i] = dhat0[i - NX];
            for (int i = 0; i < NXI * NXI; ++i) P[i] = P0[i];
            for (int i = 0; i < NU; ++i) { u[i] = u0[i]; us[i] = u0[i]; }
            int dyn_status = 0;
            std::fill(ws.begin(), ws.end(), 0.0);
            double t_start = 0.0;
            for (int k = 0; k < nsim; ++k) {
#ifdef _OPENMP
                if (k == nwarm) t_start = omp_get_wtime();
#endif
                const double t = k * h_step;
                double y[NY];
                plant_meas(x, u, t, zp, zp, y);
                if (noise) for (int i = 0; i < NY; ++i) y[i] += noise[((size_t)k * B + inst) * NY + i];
                est_update(est_type, y, u, t, zero_px, zero_py, xi, P, E);
                double d[ND + 1];
                for (int i = 0; i < ND; ++i) d[i] = (NXI > NX) ? xi[NX + i] : 0.0;
                // target problem (Target_Calc.py:41-50 parameter order; cold guess MPC_code.py:696-700)
                const double* spk = sp + (size_t)k * (NU + NY + NX);
                for (int i = 0; i < NU; ++i) parss[MPCB_OFFSS_USP + i] = spk[i];
                for (int i = 0; i < NY; ++i) parss[MPCB_OFFSS_YSP + i] = spk[NU + i];
                for (int i = 0; i < NX; ++i) parss[MPCB_OFFSS_XSP + i] = spk[NU + NY + i];
                for (int i = 0; i < ND; ++i) parss[MPCB_OFFSS_D + i] = d[i];
                for (int i = 0; i < NU; ++i) parss[MPCB_OFFSS_USPREV + i] = us[i];
                parss[MPCB_OFFSS_T] = t;
                double wss[NWS], y0[NY], fss; int st_ss, it_ss;
                for (int i = 0; i < NX; ++i) wss[i] = x0[i];
                for (int i = 0; i < NU; ++i) wss[NX + i] = u0[i];
                mdl_fy(x0, u0, d, &t, zero_py, y0);
                for (int i = 0; i < NY; ++i) wss[NZ + i] = y0[i];
                tgt_solve(parss.data(), wss, &fss, &st_ss, &it_ss, T);
                double us_prev[NU], xs_prev[NX];
                for (int i = 0; i < NU; ++i) us_prev[i] = us[i];
                for (int i = 0; i < NX; ++i) xs_prev[i] = xs[i];
                if (st_ss != 2) { for (int i = 0; i < NX; ++i) xs[i] = wss[i]; for (int i = 0; i < NU; ++i) us[i] = wss[NX + i]; }
                // OCP parameters (Control_Calc.py:43-57) and warm start (MPC_code.py:740-764)
                for (int i = 0; i < NX; ++i) { par[MPCB_OFF_X0 + i] = xi[i]; par[MPCB_OFF_XS + i] = xs[i]; }
                for (int i = 0; i < NU; ++i) { par[MPCB_OFF_US + i] = us[i]; par[MPCB_OFF_UM1 + i] = u[i]; }
                for (int i = 0; i < ND; ++i) par[MPCB_OFF_D + i] = d[i];
                par[MPCB_OFF_T] = t;
                if (k == 0) {
                    for (int i = 0; i < NW; ++i) { const int r = i % NZ; wguess[i] = (r < NX) ? x0[r] : u0[r - NX]; }
                } else if (dyn_status != 2 && dyn_status != -13) {
                    for (int i = 0; i < NW - NZ; ++i) wguess[i] = wopt[NZ + i];
                    for (int j = 0; j < NU; ++j) wguess[NW - NZ + j] = us_prev[j];
                    for (int j = 0; j < NX; ++j) wguess[NW - NX + j] = xs_prev[j];
                }
                w = wguess;
                OcpInst I = ocp_inst(ws.data(), w.data(), par.data(), &st);
                solve_ocp(I, S, scratch, max_iter_dyn);
                dyn_status = st.status;
                if (dyn_status == -13) {
                    // diverged instance: frozen
                } else if (dyn_status != 2) {
                    wopt = w;
                    for (int i = 0; i < NU; ++i) u[i] = w[NX + i];
                    for (int i = 0; i < NX; ++i) xi[i] = w[NZ + i];
                } else {
                    double xn[NX];
                    dyn_value(xi, u, d, zero_px, t, xn);
                    for (int i = 0; i < NX; ++i) xi[i] = xn[i];
                }
                for (int i = 0; i < NU; ++i) U_out[((size_t)k * B + inst) * NU + i] = u[i];
                iters_out[(size_t)k * B + inst] = st.iter; status_out[(size_t)k * B + inst] = dyn_status;
                plant_step(x, u, t, zp, zp);
            }
#ifdef _OPENMP
            if (timed_s) timed_s[inst] = (nsim > nwarm) ? omp_get_wtime() - t_start : 0.0;
#else
            (void)t_start; (void)nwarm; if (timed_s) timed_s[inst] = 0.0;
#endif
        }
    }
    return nthreads;
}

}  // extern "C"
#endif
