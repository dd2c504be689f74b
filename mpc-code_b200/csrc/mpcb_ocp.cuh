// mpcb_ocp.cuh - primal-dual interior-point solve of the dynamic OCP, stage-structured.
//
// Replaces `solver(lbx,ubx,x0,p,lbg,ubg)` at MPC_code.py:776-781 for the NLP that
// Control_Calc.py:20-260 builds (variables w=[x0,u0,...,xN], equality rows Fx(xk,uk)-x_{k+1},
// range rows Y_k in [ymin,ymax], box bounds), following IPOPT's algorithm (Waechter & Biegler
// 2006: monotone barrier, fraction-to-the-boundary, filter line search, delta_w inertia ladder).
// Instead of IPOPT's general sparse KKT factorisation the Newton system is solved by a Riccati
// recursion over the stages (x0 fixed and eliminated, range-row slacks and all bound multipliers
// condensed into the stage Hessians).
//
// Work split (one "tick" = the four kernels, each a thin wrapper around a function below):
//   ocp_eval_stage   one thread per (instance, stage): derivatives at the current iterate
//   ocp_kkt          per instance: optimality error, barrier update, Riccati solve, step sizes
//   ocp_trial_stage  one thread per (instance, stage): functions at the trial point
//   ocp_accept       per instance: filter test, accept (-> eval) or halve the step (-> trial)
// Every function is host/device so the identical code can be exercised on the CPU by the tests.
#pragma once
#include "mpcb_device.cuh"

#define NG   MPCB_NG
#define NW   MPCB_NW
#define NPAR MPCB_NPAR
#define MPCB_MAXFILT 24

enum { ST_EVAL = 0, ST_LS = 1, ST_DONE = 2 };

struct IpmOpts {       // IPOPT option names; values set from mpcb_opts_t, the rest are IPOPT defaults
    int    max_iter;
    double tol, mu_init, bound_relax, bound_push, acceptable_tol;
    int    honor_original_bounds, acceptable_iter;
};

struct InstState {
    double mu, tau, alpha, alpha_z, theta, phi, gphid, amin, theta0, dw_last, fval, E0;
    int    state, iter, status, nfilt, acc_cnt, ls_iter;
    double filt[2 * MPCB_MAXFILT];
};

// Per-instance view of the solver workspace (all device pointers).
struct OcpInst {
    double* w; const double* par;
    double *lam, *lamn, *s, *ds, *ym, *dym, *zL, *zU, *vL, *vU, *dw;
    double *A, *Bm, *c, *H, *gl, *HN, *gN, *G, *gv;
    double *Pm, *pv, *Kf, *kf, *part, *partt;
    InstState* st;
};
struct OcpShared { const double *lbx, *ubx, *lbg, *ubg; IpmOpts o; };

// ---- sizes of the per-instance workspace (doubles) -------------------------------------------
#define NGS (NG > 0 ? NG : 1)
struct OcpLayout {
    static constexpr int lam = 0;
    static constexpr int lamn = lam + NH * NX;
    static constexpr int s = lamn + NH * NX;
    static constexpr int ds = s + NH * NGS;
    static constexpr int ym = ds + NH * NGS;
    static constexpr int dym = ym + NH * NGS;
    static constexpr int vL = dym + NH * NGS;
    static constexpr int vU = vL + NH * NGS;
    static constexpr int zL = vU + NH * NGS;
    static constexpr int zU = zL + NW;
    static constexpr int dw = zU + NW;
    static constexpr int A = dw + NW;
    static constexpr int Bm = A + NH * NX * NX;
    static constexpr int c = Bm + NH * NX * NU;
    static constexpr int H = c + NH * NX;
    static constexpr int gl = H + NH * NZP;
    static constexpr int HN = gl + NH * NZ;
    static constexpr int gN = HN + NXP_;
    static constexpr int G = gN + NX;
    static constexpr int gv = G + NH * NGS * NZ;
    static constexpr int Pm = gv + NH * NGS;
    static constexpr int pv = Pm + (NH + 1) * NX * NX;
    static constexpr int Kf = pv + (NH + 1) * NX;
    static constexpr int kf = Kf + NH * NU * NX;
    static constexpr int part = kf + NH * NU;
    static constexpr int partt = part + (NH + 1) * 4;
    static constexpr int total = partt + (NH + 1) * 4;
};

MPCB_HD OcpInst ocp_inst(double* ws, double* w, const double* par, InstState* st) {
    OcpInst I;
    I.w = w; I.par = par; I.st = st;
    I.lam = ws + OcpLayout::lam; I.lamn = ws + OcpLayout::lamn; I.s = ws + OcpLayout::s; I.ds = ws + OcpLayout::ds;
    I.ym = ws + OcpLayout::ym; I.dym = ws + OcpLayout::dym; I.vL = ws + OcpLayout::vL; I.vU = ws + OcpLayout::vU;
    I.zL = ws + OcpLayout::zL; I.zU = ws + OcpLayout::zU; I.dw = ws + OcpLayout::dw;
    I.A = ws + OcpLayout::A; I.Bm = ws + OcpLayout::Bm; I.c = ws + OcpLayout::c; I.H = ws + OcpLayout::H;
    I.gl = ws + OcpLayout::gl; I.HN = ws + OcpLayout::HN; I.gN = ws + OcpLayout::gN; I.G = ws + OcpLayout::G;
    I.gv = ws + OcpLayout::gv; I.Pm = ws + OcpLayout::Pm; I.pv = ws + OcpLayout::pv; I.Kf = ws + OcpLayout::Kf;
    I.kf = ws + OcpLayout::kf; I.part = ws + OcpLayout::part; I.partt = ws + OcpLayout::partt;
    return I;
}

// ---- bounds ---------------------------------------------------------------------------------
MPCB_HD bool fin(double v) { return v > -1e300 && v < 1e300; }
MPCB_HD double rlo(double lo, double f) { return lo - f * fmax(1.0, fabs(lo)); }
MPCB_HD double rhi(double hi, double f) { return hi + f * fmax(1.0, fabs(hi)); }

// IPOPT's initial push of a primal value into its (relaxed) bounds: bound_push / bound_frac.
MPCB_HD double push_in(double v, double lo, double hi, double kappa) {
    const bool hl = fin(lo), hu = fin(hi);
    if (hl) {
        double pl = kappa * fmax(1.0, fabs(lo));
        if (hu) pl = fmin(pl, kappa * (hi - lo));
        v = fmax(v, lo + pl);
    }
    if (hu) {
        double pu = kappa * fmax(1.0, fabs(hi));
        if (hl) pu = fmin(pu, kappa * (hi - lo));
        v = fmin(v, hi - pu);
    }
    return v;
}

MPCB_HD void stage_params(const double* par, int k, double* d, double* px, double* py, double* t0) {
#pragma unroll
    for (int i = 0; i < ND; ++i) d[i] = par[MPCB_OFF_D + i];
#pragma unroll
    for (int i = 0; i < NPX; ++i) px[i] = par[MPCB_OFF_PX + k * NPX + i];
#pragma unroll
    for (int i = 0; i < NPY; ++i) py[i] = par[MPCB_OFF_PY + k * NPY + i];
    *t0 = par[MPCB_OFF_T];
}

// =============================================================================================
// init: starting point (IPOPT default initialisation), one call per (instance, stage k=0..NH)
// =============================================================================================
MPCB_HD void ocp_init_stage(OcpInst& I, const OcpShared& S, int k) {
    const double rf = S.o.bound_relax, kp = S.o.bound_push;
    double* w = I.w;
    if (k == 0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) { w[i] = I.par[MPCB_OFF_X0 + i]; I.zL[i] = 0.0; I.zU[i] = 0.0; }   // MPC_code.py:734
        InstState& st = *I.st;
        st.mu = S.o.mu_init; st.tau = fmax(0.99, 1.0 - S.o.mu_init);
        st.alpha = st.alpha_z = 0.0; st.theta0 = -1.0; st.dw_last = 0.0; st.fval = 0.0; st.E0 = 0.0;
        st.state = ST_EVAL; st.iter = 0; st.status = -1; st.nfilt = 0; st.acc_cnt = 0; st.ls_iter = 0;
    }
    const int lo_i = (k == 0) ? NX : k * NZ;
    const int hi_i = (k == NH) ? NH * NZ + NX : (k + 1) * NZ;
    for (int i = lo_i; i < hi_i; ++i) {
        const double lo = S.lbx[i], hi = S.ubx[i];
        const double lor = fin(lo) ? rlo(lo, rf) : lo, hir = fin(hi) ? rhi(hi, rf) : hi;
        w[i] = push_in(w[i], lor, hir, kp);
        I.zL[i] = fin(lo) ? 1.0 : 0.0;
        I.zU[i] = fin(hi) ? 1.0 : 0.0;
    }
    if (k < NH) {
#pragma unroll
        for (int i = 0; i < NX; ++i) I.lam[k * NX + i] = 0.0;
#if NG > 0
        double d[ND + 1], px[NPX + 1], py[NPY + 1], t0, Y[NG];
        stage_params(I.par, k, d, px, py, &t0);
        ocp_out(w + k * NZ, w + k * NZ + NX, I.par, py, Y);
        for (int i = 0; i < NG; ++i) {
            const double lo = S.lbg[k * NG + i], hi = S.ubg[k * NG + i];
            const double lor = fin(lo) ? rlo(lo, rf) : lo, hir = fin(hi) ? rhi(hi, rf) : hi;
            I.s[k * NG + i] = push_in(Y[i], lor, hir, kp);
            I.ym[k * NG + i] = 0.0;
            I.vL[k * NG + i] = fin(lo) ? 1.0 : 0.0;
            I.vU[k * NG + i] = fin(hi) ? 1.0 : 0.0;
        }
#endif
    }
}

// =============================================================================================
// eval: derivatives of stage k at the current iterate (k = 0..NH-1; k = NH-1 also does the terminal)
// =============================================================================================
MPCB_HD void ocp_eval_stage(OcpInst& I, const OcpShared& S, int k) {
    const double* w = I.w;
    double x[NX], u[NU], lam[NX], d[ND + 1], px[NPX + 1], py[NPY + 1], t0;
#pragma unroll
    for (int i = 0; i < NX; ++i) { x[i] = w[k * NZ + i]; lam[i] = I.lam[k * NX + i]; }
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = w[k * NZ + NX + i];
    stage_params(I.par, k, d, px, py, &t0);
    double xn[NX], A[NX * NX], Bm[NX * NU], Hp[NZP], l, g[NZ];
    ocp_cost_d(x, u, I.par, px, py, &l, g, Hp);          // Hp <- cost Hessian, then accumulate the rest
    dyn_full(x, u, d, px, t0, lam, xn, A, Bm, Hp);
    double th = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const double ci = xn[i] - w[(k + 1) * NZ + i];
        I.c[k * NX + i] = ci;
        th += fabs(ci);
    }
#if NG > 0
    {
        double Y[NG], JY[NG * NZ], HY[NZP], mult[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) mult[i] = I.ym[k * NG + i];
        ocp_out_d(x, u, I.par, py, mult, Y, JY, HY);
#if !MPCB_OUT_LINEAR
#pragma unroll
        for (int i = 0; i < NZP; ++i) Hp[i] += HY[i];
#endif
#pragma unroll
        for (int i = 0; i < NG; ++i) { I.gv[k * NG + i] = Y[i]; th += fabs(Y[i] - I.s[k * NG + i]); }
#pragma unroll
        for (int i = 0; i < NG * NZ; ++i) I.G[k * NG * NZ + i] = JY[i];
    }
#endif
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) I.A[k * NX * NX + i] = A[i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) I.Bm[k * NX * NU + i] = Bm[i];
#pragma unroll
    for (int i = 0; i < NZP; ++i) I.H[k * NZP + i] = Hp[i];
#pragma unroll
    for (int i = 0; i < NZ; ++i) I.gl[k * NZ + i] = g[i];
    I.part[k * 4 + 0] = l;
    I.part[k * 4 + 1] = th;
    if (k == NH - 1) {
        double V, gN[NX], HN[NXP_ + 1];
        ocp_term_d(w + NH * NZ, I.par, &V, gN, HN);
#pragma unroll
        for (int i = 0; i < NX; ++i) I.gN[i] = gN[i];
#pragma unroll
        for (int i = 0; i < NXP_; ++i) I.HN[i] = HN[i];
        I.part[NH * 4 + 0] = V;
        I.part[NH * 4 + 1] = 0.0;
    }
}

// =============================================================================================
// Lane-generic helpers.  ocp_kkt / ocp_accept are written once and run either by the 32 lanes of a
// warp (device: one warp per instance, lanes strided over stages / matrix entries, reductions by
// shuffle, small per-warp scratch in shared memory) or by a single "lane" on the host (tests).
// =============================================================================================
#ifdef __CUDA_ARCH__
#  define LANE_ID   ((int)(threadIdx.x & 31))
#  define N_LANES   32
#  define W_SYNC()  __syncwarp()
#  define W_SUM(v)  warp_sum(v)
#  define W_MAX(v)  warp_max(v)
#  define W_MIN(v)  warp_min(v)
#  define W_ISUM(v) warp_sum_int(v)
#else
#  define LANE_ID   0
#  define N_LANES   1
#  define W_SYNC()
#  define W_SUM(v)  (v)
#  define W_MAX(v)  (v)
#  define W_MIN(v)  (v)
#  define W_ISUM(v) (v)
#endif

// per-warp scratch (doubles)
struct KktScratch {
    static constexpr int P = 0;                       // NX x NX   cost-to-go Hessian of the next stage
    static constexpr int p = P + NX * NX;             // NX
    static constexpr int M = p + NX;                  // NZ x NZ   condensed stage Hessian
    static constexpr int q = M + NZ * NZ;             // NZ
    static constexpr int T = q + NZ;                  // NX x NZ   P [A B]
    static constexpr int f = T + NX * NZ;             // NX        P c + p
    static constexpr int AB = f + NX;                 // NX x NZ   [A B]
    static constexpr int K = AB + NX * NZ;            // NU x NX
    static constexpr int kk = K + NU * NX;            // NU
    static constexpr int sg = kk + NU;                // NGS       slack barrier curvature
    static constexpr int cf = sg + NGS;               // NGS       slack gradient coefficient
    static constexpr int dx = cf + NGS;               // NX
    static constexpr int du = dx + NX;                // NU
    static constexpr int dxn = du + NU;               // NX
    static constexpr int total = dxn + NX;
};

// complementarity error  max |slack * z - mu|  over all bounds
MPCB_HD double ocp_compl(const OcpInst& I, const OcpShared& S, double mu) {
    const double rf = S.o.bound_relax;
    double e = 0.0;
    for (int i = NX + LANE_ID; i < NW; i += N_LANES) {
        const double lo = S.lbx[i], hi = S.ubx[i];
        if (fin(lo)) e = fmax(e, fabs((I.w[i] - rlo(lo, rf)) * I.zL[i] - mu));
        if (fin(hi)) e = fmax(e, fabs((rhi(hi, rf) - I.w[i]) * I.zU[i] - mu));
    }
#if NG > 0
    for (int i = LANE_ID; i < NH * NG; i += N_LANES) {
        const double lo = S.lbg[i], hi = S.ubg[i];
        if (fin(lo)) e = fmax(e, fabs((I.s[i] - rlo(lo, rf)) * I.vL[i] - mu));
        if (fin(hi)) e = fmax(e, fabs((rhi(hi, rf) - I.s[i]) * I.vU[i] - mu));
    }
#endif
    return W_MAX(e);
}

struct KktErr { double dual, prim, sd, sc; };

// dual / primal infeasibility and the IPOPT scaling factors s_d, s_c (s_max = 100)
MPCB_HD KktErr ocp_errors(const OcpInst& I, const OcpShared& S) {
    double dual = 0.0, prim = 0.0, ysum = 0.0, zsum = 0.0;
    int nb = 0;
    for (int k = LANE_ID; k <= NH; k += N_LANES) {
        if (k == NH) {                                            // terminal state
            for (int j = 0; j < NX; ++j) {
                const int wi = NH * NZ + j;
                const double r = I.gN[j] - I.lam[(NH - 1) * NX + j] + I.zU[wi] - I.zL[wi];
                dual = fmax(dual, fabs(r));
                zsum += I.zL[wi] + I.zU[wi];
                nb += (fin(S.lbx[wi]) ? 1 : 0) + (fin(S.ubx[wi]) ? 1 : 0);
            }
            continue;
        }
        const double* A = I.A + k * NX * NX; const double* Bm = I.Bm + k * NX * NU;
        const double* lamn = I.lam + k * NX;                  // lam_{k+1}
        for (int j = 0; j < NZ; ++j) {
            if (k == 0 && j < NX) continue;                   // x0 is fixed
            const int wi = k * NZ + j;
            double r = I.gl[k * NZ + j];
            const double* col = (j < NX) ? (A + NX * j) : (Bm + NX * (j - NX));
            for (int i = 0; i < NX; ++i) r += col[i] * lamn[i];
            if (j < NX) r -= I.lam[(k - 1) * NX + j];         // -lam_k
#if NG > 0
            for (int i = 0; i < NG; ++i) r += I.G[k * NG * NZ + i + NG * j] * I.ym[k * NG + i];
#endif
            r += I.zU[wi] - I.zL[wi];
            dual = fmax(dual, fabs(r));
            zsum += I.zL[wi] + I.zU[wi];
            nb += (fin(S.lbx[wi]) ? 1 : 0) + (fin(S.ubx[wi]) ? 1 : 0);
        }
        for (int i = 0; i < NX; ++i) { prim = fmax(prim, fabs(I.c[k * NX + i])); ysum += fabs(lamn[i]); }
#if NG > 0
        for (int i = 0; i < NG; ++i) {
            const int gi = k * NG + i;
            prim = fmax(prim, fabs(I.gv[gi] - I.s[gi]));
            dual = fmax(dual, fabs(-I.ym[gi] - I.vL[gi] + I.vU[gi]));
            ysum += fabs(I.ym[gi]);
            zsum += I.vL[gi] + I.vU[gi];
            nb += (fin(S.lbg[gi]) ? 1 : 0) + (fin(S.ubg[gi]) ? 1 : 0);
        }
#endif
    }
    dual = W_MAX(dual); prim = W_MAX(prim); ysum = W_SUM(ysum); zsum = W_SUM(zsum); nb = W_ISUM(nb);
    const double smax = 100.0;
    const int mc = NH * NX + NH * NG;
    KktErr e;
    e.dual = dual; e.prim = prim;
    e.sd = fmax(smax, (ysum + zsum) / (double)(mc + nb > 0 ? mc + nb : 1)) / smax;
    e.sc = fmax(smax, zsum / (double)(nb > 0 ? nb : 1)) / smax;
    return e;
}

// Riccati backward sweep with regularisation dw on every primal variable.  Returns false when some
// R_k + B_k' P_{k+1} B_k is not positive definite (wrong inertia).  `sm` is the per-warp scratch.
MPCB_HD bool ocp_riccati(OcpInst& I, const OcpShared& S, double mu, double dwreg, double* sm) {
    const double rf = S.o.bound_relax;
    const int lane = LANE_ID;
    double* P = sm + KktScratch::P; double* p = sm + KktScratch::p; double* M = sm + KktScratch::M;
    double* q = sm + KktScratch::q; double* T = sm + KktScratch::T; double* f = sm + KktScratch::f;
    double* AB = sm + KktScratch::AB; double* Kk = sm + KktScratch::K; double* kk = sm + KktScratch::kk;
    double* sg = sm + KktScratch::sg; double* cf = sm + KktScratch::cf;
    // terminal stage
    for (int e = lane; e < NX * NX; e += N_LANES) {
        const int i = e % NX, j = e / NX;
        double v = I.HN[tri(i, j)];
        if (i == j) {
            const int wi = NH * NZ + i;
            const double lo = S.lbx[wi], hi = S.ubx[wi];
            v += dwreg;
            if (fin(lo)) v += I.zL[wi] / (I.w[wi] - rlo(lo, rf));
            if (fin(hi)) v += I.zU[wi] / (rhi(hi, rf) - I.w[wi]);
        }
        P[e] = v;
        I.Pm[NH * NX * NX + e] = v;
    }
    for (int i = lane; i < NX; i += N_LANES) {
        const int wi = NH * NZ + i;
        const double lo = S.lbx[wi], hi = S.ubx[wi];
        double qv = I.gN[i];
        if (fin(lo)) qv -= mu / (I.w[wi] - rlo(lo, rf));
        if (fin(hi)) qv += mu / (rhi(hi, rf) - I.w[wi]);
        p[i] = qv;
        I.pv[NH * NX + i] = qv;
    }
    W_SYNC();
    for (int k = NH - 1; k >= 0; --k) {
        // ---- (a) load [A B]; slack barrier terms of the range rows
        for (int e = lane; e < NX * NZ; e += N_LANES)
            AB[e] = (e < NX * NX) ? I.A[k * NX * NX + e] : I.Bm[k * NX * NU + (e - NX * NX)];
#if NG > 0
        for (int r = lane; r < NG; r += N_LANES) {
            const int gi = k * NG + r;
            const double lo = S.lbg[gi], hi = S.ubg[gi];
            double sig = dwreg, b = 0.0;
            if (fin(lo)) { const double dl = I.s[gi] - rlo(lo, rf); sig += I.vL[gi] / dl; b -= mu / dl; }
            if (fin(hi)) { const double du = rhi(hi, rf) - I.s[gi]; sig += I.vU[gi] / du; b += mu / du; }
            sg[r] = sig;
            cf[r] = sig * (I.gv[gi] - I.s[gi]) + b;
        }
#endif
        W_SYNC();
        // ---- (b) f = P c + p ;  T = P [A B]
        for (int i = lane; i < NX; i += N_LANES) {
            double a = p[i];
            for (int j = 0; j < NX; ++j) a += P[i + NX * j] * I.c[k * NX + j];
            f[i] = a;
        }
        for (int e = lane; e < NX * NZ; e += N_LANES) {
            const int i = e % NX, j = e / NX;
            double a = 0.0;
            for (int l = 0; l < NX; ++l) a += P[i + NX * l] * AB[l + NX * j];
            T[e] = a;
        }
        W_SYNC();
        // ---- (c) condensed stage Hessian M and gradient q, plus [A B]' P [A B] and [A B]' f
        for (int e = lane; e < NZ * NZ; e += N_LANES) {
            const int i = e % NZ, j = e / NZ;
            double m = I.H[k * NZP + tri(i, j)];
            if (i == j) {
                m += dwreg;
                if (!(k == 0 && j < NX)) {
                    const int wi = k * NZ + j;
                    const double lo = S.lbx[wi], hi = S.ubx[wi];
                    if (fin(lo)) m += I.zL[wi] / (I.w[wi] - rlo(lo, rf));
                    if (fin(hi)) m += I.zU[wi] / (rhi(hi, rf) - I.w[wi]);
                }
            }
#if NG > 0
            for (int r = 0; r < NG; ++r) m += I.G[k * NG * NZ + r + NG * i] * sg[r] * I.G[k * NG * NZ + r + NG * j];
#endif
            for (int l = 0; l < NX; ++l) m += AB[l + NX * i] * T[l + NX * j];
            M[e] = m;
        }
        for (int i = lane; i < NZ; i += N_LANES) {
            double a = I.gl[k * NZ + i];
            if (!(k == 0 && i < NX)) {
                const int wi = k * NZ + i;
                const double lo = S.lbx[wi], hi = S.ubx[wi];
                if (fin(lo)) a -= mu / (I.w[wi] - rlo(lo, rf));
                if (fin(hi)) a += mu / (rhi(hi, rf) - I.w[wi]);
            }
#if NG > 0
            for (int r = 0; r < NG; ++r) a += I.G[k * NG * NZ + r + NG * i] * cf[r];
#endif
            for (int l = 0; l < NX; ++l) a += AB[l + NX * i] * f[l];
            q[i] = a;
        }
        W_SYNC();
        // ---- (d) Cholesky of the input block M_uu = L L' (every lane, in registers)
        double L[NU * NU];
        bool pd = true;
        for (int j = 0; j < NU; ++j) {
            double djj = M[(NX + j) + NZ * (NX + j)];
            for (int l = 0; l < j; ++l) djj -= L[j + NU * l] * L[j + NU * l];
            if (!(djj > 0.0)) { pd = false; djj = 1.0; }
            djj = sqrt(djj);
            L[j + NU * j] = djj;
            for (int i = j + 1; i < NU; ++i) {
                double a = M[(NX + i) + NZ * (NX + j)];
                for (int l = 0; l < j; ++l) a -= L[i + NU * l] * L[j + NU * l];
                L[i + NU * j] = a / djj;
            }
        }
        if (!pd) return false;                                   // uniform across the warp
        // ---- (e) K = -Muu^{-1} Mux (NU x NX), kff = -Muu^{-1} q_u : one column per lane
        for (int c = lane; c <= NX; c += N_LANES) {
            double y[NU];
            for (int i = 0; i < NU; ++i) {
                double a = (c < NX) ? M[(NX + i) + NZ * c] : q[NX + i];
                for (int l = 0; l < i; ++l) a -= L[i + NU * l] * y[l];
                y[i] = a / L[i + NU * i];
            }
            for (int i = NU - 1; i >= 0; --i) {
                double a = y[i];
                for (int l = i + 1; l < NU; ++l) a -= L[l + NU * i] * y[l];
                y[i] = a / L[i + NU * i];
            }
            for (int i = 0; i < NU; ++i) {
                if (c < NX) { Kk[i + NU * c] = -y[i]; I.Kf[k * NU * NX + i + NU * c] = -y[i]; }
                else { kk[i] = -y[i]; I.kf[k * NU + i] = -y[i]; }
            }
        }
        W_SYNC();
        // ---- (f) P = Mxx + Mxu K (symmetrised), p = q_x + Mxu kff
        for (int e = lane; e < NX * NX; e += N_LANES) {
            const int i = e % NX, j = e / NX;
            double a = M[i + NZ * j], b = M[j + NZ * i];
            for (int l = 0; l < NU; ++l) { a += M[i + NZ * (NX + l)] * Kk[l + NU * j]; b += M[j + NZ * (NX + l)] * Kk[l + NU * i]; }
            const double v = (i == j) ? a : 0.5 * (a + b);
            P[e] = v;
            I.Pm[k * NX * NX + e] = v;
        }
        for (int i = lane; i < NX; i += N_LANES) {
            double a = q[i];
            for (int l = 0; l < NU; ++l) a += M[i + NZ * (NX + l)] * kk[l];
            p[i] = a;
            I.pv[k * NX + i] = a;
        }
        W_SYNC();
    }
    return true;
}

MPCB_HD void ocp_finish(OcpInst& I, const OcpShared& S, int status) {
    InstState& st = *I.st;
    double f = 0.0;
    for (int k = LANE_ID; k <= NH; k += N_LANES) f += I.part[k * 4 + 0];
    f = W_SUM(f);
    if (S.o.honor_original_bounds)
        for (int i = NX + LANE_ID; i < NW; i += N_LANES) I.w[i] = fmin(fmax(I.w[i], S.lbx[i]), S.ubx[i]);
    if (LANE_ID == 0) { st.status = status; st.state = ST_DONE; st.fval = f; }
    W_SYNC();
}

// =============================================================================================
// kkt: one interior-point iteration up to (not including) the line search, for one instance
// =============================================================================================
MPCB_HD void ocp_kkt(OcpInst& I, const OcpShared& S, double* sm) {
    InstState& st = *I.st;
    const double rf = S.o.bound_relax;
    const int lane = LANE_ID;
    const int iter = st.iter;
#if NG > 0
    // A stage-0 range row that does not depend on u_0 is a constant (x_0 is fixed).  Outside its relaxed
    // bounds the OCP is infeasible: IPOPT would end in restoration with Infeasible_Problem_Detected, the
    // one status the reference loop reacts to (MPC_code.py:786,804; quirk D7 of SURVEY.md).
    if (iter == 0) {
        bool infeasible = false;
        for (int r = 0; r < NG; ++r) {
            bool constant = true;
            for (int j = NX; j < NZ; ++j) if (I.G[r + NG * j] != 0.0) constant = false;
            if (!constant) continue;
            const double v = I.gv[r], lo = S.lbg[r], hi = S.ubg[r];
            if ((fin(lo) && v < rlo(lo, rf) - S.o.tol) || (fin(hi) && v > rhi(hi, rf) + S.o.tol)) infeasible = true;
        }
        if (infeasible) { ocp_finish(I, S, 2); return; }
    }
#endif
    // ---- optimality error and termination (IPOPT: tol, dual_inf_tol=1, constr_viol_tol=1e-4, compl_inf_tol=1e-4)
    const KktErr e = ocp_errors(I, S);
    const double c0 = ocp_compl(I, S, 0.0);
    const double E0 = fmax(fmax(e.dual / e.sd, e.prim), c0 / e.sc);
    int acc_cnt = st.acc_cnt;
    W_SYNC();
    if (lane == 0) st.E0 = E0;
    if (!(E0 == E0) || !fin(E0)) { ocp_finish(I, S, -13); return; }
    if (E0 <= S.o.tol && e.dual <= 1.0 && e.prim <= 1e-4 && c0 <= 1e-4) { ocp_finish(I, S, 0); return; }
    if (E0 <= S.o.acceptable_tol && e.dual <= 1e10 && e.prim <= 1e-2 && c0 <= 1e-2) {
        acc_cnt += 1;
        if (acc_cnt >= S.o.acceptable_iter) { ocp_finish(I, S, 1); return; }
    } else {
        acc_cnt = 0;
    }
    if (iter >= S.o.max_iter) { ocp_finish(I, S, -1); return; }
    // ---- monotone barrier update (kappa_eps=10, kappa_mu=0.2, theta_mu=1.5)
    double mu = st.mu;
    const double mu_min = S.o.tol / 10.0;
    bool changed = false;
    while (mu > mu_min) {
        const double Emu = fmax(fmax(e.dual / e.sd, e.prim), ocp_compl(I, S, mu) / e.sc);
        if (Emu > 10.0 * mu) break;
        mu = fmax(mu_min, fmin(0.2 * mu, pow(mu, 1.5)));
        changed = true;
    }
    const double tau = changed ? fmax(0.99, 1.0 - mu) : st.tau;
    const double dw_last = st.dw_last;
    const double theta0_old = st.theta0;
    W_SYNC();
    // ---- Newton step by Riccati recursion, inertia-correcting ladder on delta_w
    double dwreg = 0.0;
    bool first = true, ok = false;
    for (int attempt = 0; attempt < 60; ++attempt) {
        if (ocp_riccati(I, S, mu, dwreg, sm)) { ok = true; break; }
        W_SYNC();
        if (first) { dwreg = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0); first = false; }
        else dwreg *= (dw_last == 0.0) ? 100.0 : 8.0;
        if (dwreg > 1e40) break;
    }
    if (!ok) { ocp_finish(I, S, -3); return; }
    // ---- forward sweep (sequential over the stages): dw and the new dynamics multipliers
    double* dx = sm + KktScratch::dx; double* du = sm + KktScratch::du; double* dxn = sm + KktScratch::dxn;
    for (int i = lane; i < NX; i += N_LANES) { dx[i] = 0.0; I.dw[i] = 0.0; }
    W_SYNC();
    for (int k = 0; k < NH; ++k) {
        const double* A = I.A + k * NX * NX; const double* Bm = I.Bm + k * NX * NU;
        for (int i = lane; i < NU; i += N_LANES) {
            double a = I.kf[k * NU + i];
            for (int j = 0; j < NX; ++j) a += I.Kf[k * NU * NX + i + NU * j] * dx[j];
            du[i] = a;
            I.dw[k * NZ + NX + i] = a;
        }
        W_SYNC();
        for (int i = lane; i < NX; i += N_LANES) {
            double a = I.c[k * NX + i];
            for (int j = 0; j < NX; ++j) a += A[i + NX * j] * dx[j];
            for (int j = 0; j < NU; ++j) a += Bm[i + NX * j] * du[j];
            dxn[i] = a;
            I.dw[(k + 1) * NZ + i] = a;
        }
        W_SYNC();
        for (int i = lane; i < NX; i += N_LANES) {
            double a = I.pv[(k + 1) * NX + i];
            for (int j = 0; j < NX; ++j) a += I.Pm[(k + 1) * NX * NX + i + NX * j] * dxn[j];
            I.lamn[k * NX + i] = a;
            dx[i] = dxn[i];
        }
        W_SYNC();
    }
    // ---- stage-parallel: slack / multiplier steps of the range rows, fraction to the boundary (primal and
    //      dual), barrier objective and its directional derivative
    double amax = 1.0, az = 1.0, gphid = 0.0, theta = 0.0, barr = 0.0, fobj = 0.0;
    for (int k = lane; k <= NH; k += N_LANES) {
        fobj += I.part[k * 4 + 0];
        if (k < NH) theta += I.part[k * 4 + 1];
#if NG > 0
        if (k < NH) {
            for (int r = 0; r < NG; ++r) {
                const int gi = k * NG + r;
                const double lo = S.lbg[gi], hi = S.ubg[gi];
                double sig = dwreg, b = 0.0, dl = 1.0, du_ = 1.0;
                if (fin(lo)) { dl = I.s[gi] - rlo(lo, rf); sig += I.vL[gi] / dl; b -= mu / dl; barr += log(dl); }
                if (fin(hi)) { du_ = rhi(hi, rf) - I.s[gi]; sig += I.vU[gi] / du_; b += mu / du_; barr += log(du_); }
                double dsr = I.gv[gi] - I.s[gi];
                for (int j = 0; j < NZ; ++j) dsr += I.G[k * NG * NZ + r + NG * j] * I.dw[k * NZ + j];
                I.ds[gi] = dsr;
                I.dym[gi] = sig * dsr + b - I.ym[gi];
                gphid += b * dsr;
                if (fin(lo)) {
                    if (dsr < 0.0) amax = fmin(amax, -tau * dl / dsr);
                    const double dz = mu / dl - I.vL[gi] - I.vL[gi] / dl * dsr;
                    if (dz < 0.0) az = fmin(az, -tau * I.vL[gi] / dz);
                }
                if (fin(hi)) {
                    if (dsr > 0.0) amax = fmin(amax, tau * du_ / dsr);
                    const double dz = mu / du_ - I.vU[gi] + I.vU[gi] / du_ * dsr;
                    if (dz < 0.0) az = fmin(az, -tau * I.vU[gi] / dz);
                }
            }
        }
#endif
        const int nz = (k == NH) ? NX : NZ;
        for (int j = 0; j < nz; ++j) {
            if (k == 0 && j < NX) continue;
            const int wi = k * NZ + j;
            const double dv = I.dw[wi];
            gphid += ((k == NH) ? I.gN[j] : I.gl[k * NZ + j]) * dv;
            const double lo = S.lbx[wi], hi = S.ubx[wi];
            if (fin(lo)) {
                const double dl = I.w[wi] - rlo(lo, rf);
                barr += log(dl); gphid -= mu / dl * dv;
                if (dv < 0.0) amax = fmin(amax, -tau * dl / dv);
                const double dz = mu / dl - I.zL[wi] - I.zL[wi] / dl * dv;
                if (dz < 0.0) az = fmin(az, -tau * I.zL[wi] / dz);
            }
            if (fin(hi)) {
                const double du_ = rhi(hi, rf) - I.w[wi];
                barr += log(du_); gphid += mu / du_ * dv;
                if (dv > 0.0) amax = fmin(amax, tau * du_ / dv);
                const double dz = mu / du_ - I.zU[wi] + I.zU[wi] / du_ * dv;
                if (dz < 0.0) az = fmin(az, -tau * I.zU[wi] / dz);
            }
        }
    }
    amax = W_MIN(amax); az = W_MIN(az); gphid = W_SUM(gphid); theta = W_SUM(theta); barr = W_SUM(barr); fobj = W_SUM(fobj);
    const double phi = fobj - mu * barr;
    const double theta0 = (theta0_old < 0.0) ? theta : theta0_old;
    const double theta_min = 1e-4 * fmax(1.0, theta0);
    // minimal step size before the line search gives up (gamma_alpha=0.05, gamma_theta=1e-5, gamma_phi=1e-8)
    double amin;
    if (gphid < 0.0 && theta <= theta_min) {
        amin = fmin(1e-5, fmin(1e-8 * theta / (-gphid), pow(theta, 1.1) / pow(-gphid, 2.3)));
    } else if (gphid < 0.0) {
        amin = fmin(1e-5, 1e-8 * theta / (-gphid));
    } else {
        amin = 1e-5;
    }
    if (lane == 0) {
        st.acc_cnt = acc_cnt;
        if (changed) { st.mu = mu; st.tau = tau; st.nfilt = 0; }
        if (dwreg > 0.0) st.dw_last = dwreg;
        st.theta0 = theta0;
        st.amin = 0.05 * amin;
        st.theta = theta; st.phi = phi; st.gphid = gphid;
        st.alpha = amax; st.alpha_z = az;
        st.ls_iter = 0;
        st.state = ST_LS;
    }
    W_SYNC();
}

// =============================================================================================
// trial: constraint violation, objective and barrier terms of stage k at w + alpha dw
// =============================================================================================
MPCB_HD void ocp_trial_stage(OcpInst& I, const OcpShared& S, int k) {
    const double al = I.st->alpha, rf = S.o.bound_relax;
    const double* w = I.w; const double* dw = I.dw;
    double x[NX], u[NU], d[ND + 1], px[NPX + 1], py[NPY + 1], t0, xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = w[k * NZ + i] + al * dw[k * NZ + i];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = w[k * NZ + NX + i] + al * dw[k * NZ + NX + i];
    stage_params(I.par, k, d, px, py, &t0);
    dyn_value(x, u, d, px, t0, xn);
    double th = 0.0, barr = 0.0, l;
#pragma unroll
    for (int i = 0; i < NX; ++i) th += fabs(xn[i] - (w[(k + 1) * NZ + i] + al * dw[(k + 1) * NZ + i]));
    ocp_cost(x, u, I.par, px, py, &l);
#if NG > 0
    {
        double Y[NG];
        ocp_out(x, u, I.par, py, Y);
        for (int i = 0; i < NG; ++i) {
            const int gi = k * NG + i;
            const double st_ = I.s[gi] + al * I.ds[gi];
            th += fabs(Y[i] - st_);
            const double lo = S.lbg[gi], hi = S.ubg[gi];
            if (fin(lo)) barr += log(st_ - rlo(lo, rf));
            if (fin(hi)) barr += log(rhi(hi, rf) - st_);
        }
    }
#endif
    for (int j = (k == 0 ? NX : 0); j < NZ; ++j) {
        const int wi = k * NZ + j;
        const double v = (j < NX) ? x[j] : u[j - NX];
        const double lo = S.lbx[wi], hi = S.ubx[wi];
        if (fin(lo)) barr += log(v - rlo(lo, rf));
        if (fin(hi)) barr += log(rhi(hi, rf) - v);
    }
    I.partt[k * 4 + 0] = l; I.partt[k * 4 + 1] = th; I.partt[k * 4 + 2] = barr;
    if (k == NH - 1) {
        double xN[NX], V, bN = 0.0;
        for (int j = 0; j < NX; ++j) {
            const int wi = NH * NZ + j;
            xN[j] = w[wi] + al * dw[wi];
            const double lo = S.lbx[wi], hi = S.ubx[wi];
            if (fin(lo)) bN += log(xN[j] - rlo(lo, rf));
            if (fin(hi)) bN += log(rhi(hi, rf) - xN[j]);
        }
        ocp_term(xN, I.par, &V);
        I.partt[NH * 4 + 0] = V; I.partt[NH * 4 + 1] = 0.0; I.partt[NH * 4 + 2] = bN;
    }
}

// =============================================================================================
// accept: filter line-search decision for one instance (lane-generic, see above)
// =============================================================================================
MPCB_HD void ocp_accept(OcpInst& I, const OcpShared& S) {
    InstState& st = *I.st;
    const int lane = LANE_ID;
    const double rf = S.o.bound_relax, mu = st.mu;
    double th_t = 0.0, f_t = 0.0, b_t = 0.0;
    for (int k = lane; k <= NH; k += N_LANES) { f_t += I.partt[k * 4 + 0]; th_t += I.partt[k * 4 + 1]; b_t += I.partt[k * 4 + 2]; }
    th_t = W_SUM(th_t); f_t = W_SUM(f_t); b_t = W_SUM(b_t);
    const double ph_t = f_t - mu * b_t;
    const double theta = st.theta, phi = st.phi, gphid = st.gphid, alpha = st.alpha, az = st.alpha_z;
    const double theta_min = 1e-4 * fmax(1.0, st.theta0), theta_max = 1e4 * fmax(1.0, st.theta0);
    const int nfilt = st.nfilt;
    bool ok = (th_t == th_t) && (ph_t == ph_t) && fin(th_t) && fin(ph_t) && th_t <= theta_max;
    if (ok)
        for (int i = 0; i < nfilt; ++i)
            if (th_t >= st.filt[2 * i] && ph_t >= st.filt[2 * i + 1]) { ok = false; break; }
    bool accepted = false, ftype = false;
    if (ok) {
        const bool switching = gphid < 0.0 && theta <= theta_min && alpha * pow(-gphid, 2.3) > pow(theta, 1.1);
        const double eps = 10.0 * 2.220446049250313e-16 * fabs(phi);
        if (switching) {
            if (ph_t - phi - eps <= 1e-8 * alpha * gphid) { accepted = true; ftype = true; }
        } else {
            if (th_t <= (1.0 - 1e-5) * theta || ph_t - eps <= phi - 1e-8 * theta) accepted = true;
        }
    }
    const double amin = st.amin;
    W_SYNC();
    if (!accepted) {
        const double an = 0.5 * alpha;
        const bool give_up = !(an >= amin * (1.0 - 1e-12)) || an <= 1e-16;
        if (lane == 0) { st.alpha = an; st.ls_iter += 1; }
        W_SYNC();
        // No restoration phase.  IPOPT would now minimise the constraint violation; when that cannot be reduced it
        // returns Infeasible_Problem_Detected (the status the reference loop acts on), so a failed line search away
        // from feasibility (violation above constr_viol_tol = 1e-4) is reported as 2, otherwise Restoration_Failed.
        if (give_up) ocp_finish(I, S, theta > 1e-4 ? 2 : -2);
        return;
    }
    // ---- take the step; bound multipliers with their own step size, then the kappa_sigma safeguard
    const double ks = 1e10;
    for (int i = NX + lane; i < NW; i += N_LANES) {
        const double lo = S.lbx[i], hi = S.ubx[i], dv = I.dw[i];
        const double wn = I.w[i] + alpha * dv;
        if (fin(lo)) {
            const double dl = I.w[i] - rlo(lo, rf), dln = wn - rlo(lo, rf);
            double z = I.zL[i] + az * (mu / dl - I.zL[i] - I.zL[i] / dl * dv);
            I.zL[i] = fmax(fmin(z, ks * mu / dln), mu / (ks * dln));
        }
        if (fin(hi)) {
            const double du = rhi(hi, rf) - I.w[i], dun = rhi(hi, rf) - wn;
            double z = I.zU[i] + az * (mu / du - I.zU[i] + I.zU[i] / du * dv);
            I.zU[i] = fmax(fmin(z, ks * mu / dun), mu / (ks * dun));
        }
        I.w[i] = wn;
    }
    for (int i = lane; i < NH * NX; i += N_LANES) I.lam[i] += alpha * (I.lamn[i] - I.lam[i]);
#if NG > 0
    for (int i = lane; i < NH * NG; i += N_LANES) {
        const double lo = S.lbg[i], hi = S.ubg[i], dv = I.ds[i];
        const double sn = I.s[i] + alpha * dv;
        if (fin(lo)) {
            const double dl = I.s[i] - rlo(lo, rf), dln = sn - rlo(lo, rf);
            double z = I.vL[i] + az * (mu / dl - I.vL[i] - I.vL[i] / dl * dv);
            I.vL[i] = fmax(fmin(z, ks * mu / dln), mu / (ks * dln));
        }
        if (fin(hi)) {
            const double du = rhi(hi, rf) - I.s[i], dun = rhi(hi, rf) - sn;
            double z = I.vU[i] + az * (mu / du - I.vU[i] + I.vU[i] / du * dv);
            I.vU[i] = fmax(fmin(z, ks * mu / dun), mu / (ks * dun));
        }
        I.s[i] = sn;
        I.ym[i] += alpha * I.dym[i];
    }
#endif
    if (lane == 0) {
        if (!ftype && nfilt < MPCB_MAXFILT) {
            st.filt[2 * nfilt] = (1.0 - 1e-5) * theta;
            st.filt[2 * nfilt + 1] = phi - 1e-8 * theta;
            st.nfilt = nfilt + 1;
        }
        st.iter += 1;
        st.state = ST_EVAL;
    }
    W_SYNC();
}
