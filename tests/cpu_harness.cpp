// cpu_harness.cpp - TEST INFRASTRUCTURE: runs the host/device functions of mpc-code_b200/csrc/*.cuh
// on the CPU (compiled with g++, no CUDA), with the same tick schedule as the kernel driver in
// mpcb_api.cu.  It lets the `-m "not gpu"` tests exercise the device arithmetic (RK4 sweeps,
// Riccati interior-point iteration, target solve, estimator) against the oracle without a GPU.
// It is NOT part of the product: nothing in mpc-code_b200/ loads it.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "mpcb_target.cuh"

static IpmOpts mk_opts(int max_iter, double tol, double mu_init, double relax, int honor) {
    IpmOpts o; o.max_iter = max_iter; o.tol = tol; o.mu_init = mu_init; o.bound_relax = relax; o.bound_push = 1e-2;
    o.acceptable_tol = 1e-6; o.honor_original_bounds = honor; o.acceptable_iter = 15; return o;
}

extern "C" {

void h_dims(int* out) {
    int v[] = {NX, NU, NY, ND, NPX, NPY, NXI, NH, MX,
#if MPCB_HAS_OCP
               NW, NPAR, NG,
#else
               0, 0, 0,
#endif
#if MPCB_HAS_TARGET
               NWS, MPCB_NPARSS
#else
               0, 0
#endif
    };
    memcpy(out, v, sizeof(v));
}

#if MPCB_HAS_OCP
void h_stage_derivs(int B, const double* par, const double* w, const double* lam, double* A, double* Bm, double* c, double* H) {
    for (int inst = 0; inst < B; ++inst) for (int k = 0; k < NH; ++k) {
        const double* wi = w + (size_t)inst * NW; const double* pi = par + (size_t)inst * NPAR;
        double x[NX], u[NU], l[NX], d[ND + 1], px[NPX + 1], py[NPY + 1], t0;
        for (int i = 0; i < NX; ++i) { x[i] = wi[k * NZ + i]; l[i] = lam[((size_t)inst * NH + k) * NX + i]; }
        for (int i = 0; i < NU; ++i) u[i] = wi[k * NZ + NX + i];
        stage_params(pi, k, d, px, py, &t0);
        double xn[NX], Al[NX * NX], Bl[NX * NU], Hp[NZP];
        for (int i = 0; i < NZP; ++i) Hp[i] = 0.0;
        dyn_full(x, u, d, px, t0, l, xn, Al, Bl, Hp);
        const size_t s = (size_t)inst * NH + k;
        for (int i = 0; i < NX * NX; ++i) A[s * NX * NX + i] = Al[i];
        for (int i = 0; i < NX * NU; ++i) Bm[s * NX * NU + i] = Bl[i];
        for (int i = 0; i < NX; ++i) c[s * NX + i] = xn[i] - wi[(k + 1) * NZ + i];
        for (int i = 0; i < NZP; ++i) H[s * NZP + i] = Hp[i];
        // first-order-only sweep must agree with the full sweep
        double xn2[NX], A2[NX * NX], B2[NX * NU];
        dyn_sens(x, u, d, px, t0, xn2, A2, B2);
        for (int i = 0; i < NX * NX; ++i) if (fabs(A2[i] - Al[i]) > 1e-12 * (1 + fabs(Al[i]))) abort();
    }
}

int h_ocp(int B, const double* par, double* w, double* f, int* status, int* iters,
          const double* lbx, const double* ubx, const double* lbg, const double* ubg,
          int max_iter, double tol, double mu_init, double relax, int honor, int* ticks_out) {
    // bounds: reference layout -> internal layout (as mpcb_set_const does)
    std::vector<double> lbi(NWI, -INFINITY), ubi(NWI, INFINITY);
    for (int k = 0; k <= NH; ++k) {
        for (int i = 0; i < NX; ++i) { lbi[k * NZA + i] = lbx[k * NZ + i]; ubi[k * NZA + i] = ubx[k * NZ + i]; }
        if (k < NH) for (int i = 0; i < NU; ++i) { lbi[k * NZA + NXA + i] = lbx[k * NZ + NX + i]; ubi[k * NZA + NXA + i] = ubx[k * NZ + NX + i]; }
    }
    OcpShared S; S.lbx = lbi.data(); S.ubx = ubi.data(); S.lbg = lbg; S.ubg = ubg; S.o = mk_opts(max_iter, tol, mu_init, relax, honor);
    std::vector<double> ws((size_t)B * OcpLayout::total, 0.0);
    std::vector<InstState> st(B);
    auto view = [&](int inst) { return ocp_inst(ws.data() + (size_t)inst * OcpLayout::total, w + (size_t)inst * NW,
                                                par + (size_t)inst * NPAR, &st[inst]); };
    for (int inst = 0; inst < B; ++inst) for (int k = 0; k <= NH; ++k) { OcpInst I = view(inst); ocp_init_stage(I, S, k); }
    int ticks = 0;
    const int max_ticks = (max_iter + 2) * 8;
    while (ticks < max_ticks) {
        int active = 0;
        for (int inst = 0; inst < B; ++inst) {
            OcpInst I = view(inst);
            double scratch[KktScratch::total];
            if (st[inst].state == ST_EVAL) { for (int k = 0; k < NH; ++k) ocp_eval_stage(I, S, k); ocp_kkt(I, S, scratch); }
            if (st[inst].state == ST_LS) { for (int k = 0; k < NH; ++k) ocp_trial_stage(I, S, k); ocp_accept(I, S); }
            if (st[inst].state != ST_DONE) active++;
        }
        ticks++;
        if (!active) break;
    }
    for (int inst = 0; inst < B; ++inst) {
        OcpInst I = view(inst);
        for (int k = 0; k <= NH; ++k) ocp_export_stage(I, k);
        f[inst] = st[inst].fval; status[inst] = st[inst].status; iters[inst] = st[inst].iter;
    }
    if (ticks_out) *ticks_out = ticks;
    return 0;
}
#endif

#if MPCB_HAS_TARGET
int h_target(int B, const double* par, double* w, double* f, int* status, int* iters, const double* lbx, const double* ubx,
             int max_iter, double tol, double mu_init, double relax, int honor) {
    TgtShared S; S.lbx = lbx; S.ubx = ubx; S.o = mk_opts(max_iter, tol, mu_init, relax, honor);
    for (int inst = 0; inst < B; ++inst)
        tgt_solve(par + (size_t)inst * MPCB_NPARSS, w + (size_t)inst * NWS, f + inst, status + inst, iters + inst, S);
    return 0;
}
#endif

void h_estimate(int B, int est_type, const double* y, const double* u, const double* t, const double* px, const double* py,
                double* xi, double* P, const double* Q, const double* R, const double* K, const double* dmin, const double* dmax,
                int has_dbounds) {
    EstShared E; E.Q = Q; E.R = R; E.K = K; E.dmin = dmin; E.dmax = dmax; E.has_dbounds = has_dbounds;
    for (int inst = 0; inst < B; ++inst)
        est_update(est_type, y + (size_t)inst * NY, u + (size_t)inst * NU, t[inst], px + (size_t)inst * NPX,
                   py + (size_t)inst * NPY, xi + (size_t)inst * NXI, P + (size_t)inst * NXI * NXI, E);
}

void h_model_step(int B, const double* x, const double* u, const double* d, const double* t, const double* px, double* xn) {
    for (int inst = 0; inst < B; ++inst) {
        double dl[ND + 1], pl[NPX + 1];
        for (int i = 0; i < ND; ++i) dl[i] = d[(size_t)inst * ND + i];
        for (int i = 0; i < NPX; ++i) pl[i] = px[(size_t)inst * NPX + i];
        dyn_value(x + (size_t)inst * NX, u + (size_t)inst * NU, dl, pl, t[inst], xn + (size_t)inst * NX);
    }
}

#if !MPCB_PLANT_NOMINAL
void h_plant_step(int B, double* x, const double* u, const double* t, const double* pxp, const double* pxmp) {
    for (int inst = 0; inst < B; ++inst)
        plant_step(x + (size_t)inst * MPCB_NXP, u + (size_t)inst * NU, t[inst], pxp + (size_t)inst * MPCB_NPXP,
                   pxmp + (size_t)inst * MPCB_NPXP);
}
void h_plant_meas(int B, const double* x, const double* u, const double* t, const double* pyp, const double* pymp, double* y) {
    for (int inst = 0; inst < B; ++inst)
        plant_meas(x + (size_t)inst * MPCB_NXP, u + (size_t)inst * NU, t[inst], pyp + (size_t)inst * MPCB_NPYP,
                   pymp + (size_t)inst * MPCB_NPYP, y + (size_t)inst * NY);
}
#endif

void h_bk_test(int n_unused, double* A12, double* b12, int* inertia) {
    int ipiv[12], np_, nn_, nz_;
    bk_factor<12>(A12, ipiv, &np_, &nn_, &nz_);
    bk_solve<12>(A12, ipiv, b12);
    inertia[0] = np_; inertia[1] = nn_; inertia[2] = nz_;
}

}  // extern "C"
