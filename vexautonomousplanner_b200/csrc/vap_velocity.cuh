// vap_velocity.cuh -- v2 of stages S3-events / S4 / S5: sample-parallel event detection, a pre-pass that hoists
// every state-independent term of the forward / backward recurrences into one 32-byte record per sample, and
// chunk-speculative kernels that run the exact serial recurrences on many chunks of one path at once.
//
// Exactness: the recurrences of motion_profile_generator.py:188-311 are NOT associative (the wheel-acceleration
// term depends on v[i] and v[i-1]), so no scan is used.  A chunk starts from a guessed state, and
// is re-run from its predecessor's true end state until two consecutive velocities are BITWISE equal to the ones
// computed before; from there on the old results are the serial results.  The fix-up loop ends when no chunk
// changed, at which point (by induction from chunk 0) every value equals the serial evaluation bit for bit.
#pragma once
#include "vap_device.cuh"

#define EV_AP_CAND 4

// read-only (ld.global.nc) load of a 32-byte record
__device__ __forceinline__ double4 ldg_d4(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ---- S3 (parallel): t, kappa, theta per distance sample + event candidates -----------------------------------
// wrap candidates: samples with frac(t[i-1]) > frac(t[i]) and t[i] < N-1  (motion_profile_generator.py:124)
// action candidates: samples with t[i-1] < ap.t <= t[i]                   (:142-146)
__global__ void __launch_bounds__(256) k_dist_sample_ev(
    int N_max, int A_max, const int* __restrict__ n_nodes, const int* __restrict__ n_splines,
    const int* __restrict__ status, const double* __restrict__ ap_attr, const int* __restrict__ n_ap,
    const double* __restrict__ dgrid, int samples, long long Q_cap, const double* __restrict__ lut_d,
    const double* __restrict__ lut_t, const double* __restrict__ total_len, int spn, long long P_cap,
    const double* __restrict__ prop_k, const double* __restrict__ prop_h, long long D_cap,
    const int* __restrict__ n_samples, double* __restrict__ t_out, double* __restrict__ kap, double* __restrict__ th,
    int* __restrict__ ev_wrap, int* __restrict__ ev_nwrap, int* __restrict__ ev_apc, int* __restrict__ ev_napc)
{
    __shared__ double s_t[257];
    long long b = blockIdx.y;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int i0 = blockIdx.x * blockDim.x;
    if (i0 >= D) return;
    const int i = i0 + threadIdx.x;
    const int n = n_nodes[b];
    const double* ld = lut_d + (size_t)b * Q_cap;
    const double* lt = lut_t + (size_t)b * Q_cap;
    const int Q = samples * n_splines[b];
    const double L = total_len[b];
    const PropGrid pg = prop_grid(spn, n);
    double t = 0.0;
    if (i < D) {
        t = (i == D - 1) ? (double)(n - 1) : distance_to_time32(ld, lt, Q, L, n, dgrid[i]);
        double k, h;
        snap_gather2_32(prop_k + (size_t)b * P_cap, prop_h + (size_t)b * P_cap, t, pg, k, h);
        size_t o = (size_t)b * D_cap + i;
        t_out[o] = t; kap[o] = k; th[o] = h;
    }
    s_t[threadIdx.x + 1] = t;
    if (threadIdx.x == 0) s_t[0] = (i0 > 0) ? distance_to_time32(ld, lt, Q, L, n, dgrid[i0 - 1]) : 0.0;   // prev_t of sample 0 is 0
    __syncthreads();
    if (i >= D - 1) return;            // the final appended sample takes no part in the event logic
    double tp = s_t[threadIdx.x];
    if (frac1(tp) > frac1(t) && t < (double)(n - 1)) {
        int slot = atomicAdd(ev_nwrap + b, 1);
        if (slot < N_max) ev_wrap[(size_t)b * N_max + slot] = (int)i;
    }
    int A = n_ap ? n_ap[b] : 0;
    for (int k = 0; k < A; k++) {
        double x = ap_attr[((size_t)b * A_max + k) * APA + P_T];
        if (tp < x && t >= x) {
            int slot = atomicAdd(ev_napc + (size_t)b * A_max + k, 1);
            if (slot < EV_AP_CAND) ev_apc[((size_t)b * A_max + k) * EV_AP_CAND + slot] = (int)i;
        }
    }
}

// ---- S3 (per path): replay the sampling loop's event logic over the candidate samples only --------------------
// Outputs the reference's max_accels / boundary_map plus the piecewise-constant initial-velocity regimes:
//   vr_idx/vr_val: v0[i] = vr_val[j] for the last j with vr_idx[j] <= i      (max_velocity before sample i's events)
//   st_idx       : samples whose initial velocity is overwritten with 0.01    (stop nodes / stop action points)
__global__ void k_resolve_events(long long B, int N_max, int A_max, const double* __restrict__ node_attr,
                                 const int* __restrict__ node_flags, const int* __restrict__ n_nodes,
                                 const double* __restrict__ ap_attr, const int* __restrict__ ap_flags,
                                 const int* __restrict__ n_ap, const double* __restrict__ cons,
                                 int* __restrict__ status, int* __restrict__ ev_wrap, const int* __restrict__ ev_nwrap,
                                 const int* __restrict__ ev_apc, const int* __restrict__ ev_napc, int E_cap,
                                 double* __restrict__ max_accels, int* __restrict__ bidx, int* __restrict__ bval,
                                 int* __restrict__ n_ev, int* __restrict__ vr_idx, double* __restrict__ vr_val,
                                 int* __restrict__ st_idx, int* __restrict__ n_vr, double dt,
                                 float* __restrict__ ins_est)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    n_ev[2 * b] = 0; n_ev[2 * b + 1] = 0; n_vr[2 * b] = 0; n_vr[2 * b + 1] = 0;
    if (ins_est) ins_est[b] = 0.f;
    if (status[b] != ST_OK) return;
    const int n = n_nodes[b];
    const int A = n_ap ? n_ap[b] : 0;
    const double* na = node_attr + (size_t)b * N_max * NA;
    const int* nf = node_flags + (size_t)b * N_max;
    const double* apa = ap_attr + (size_t)b * A_max * APA;
    const int* apf = ap_flags + (size_t)b * A_max;
    const double V = cons[b * 6 + 0], A0 = cons[b * 6 + 1];
    double* ma = max_accels + (size_t)b * E_cap;
    int* bi = bidx + (size_t)b * E_cap;
    int* bv = bval + (size_t)b * E_cap;
    int* vi = vr_idx + (size_t)b * E_cap;
    double* vv = vr_val + (size_t)b * E_cap;
    int* si = st_idx + (size_t)b * E_cap;
    int* wr = ev_wrap + (size_t)b * N_max;
    int nw = ev_nwrap[b];
    if (nw > N_max) { status[b] = ST_CAPACITY; return; }
    for (int i = 1; i < nw; i++) {           // sort wrap samples ascending
        int x = wr[i], j = i - 1;
        while (j >= 0 && wr[j] > x) { wr[j + 1] = wr[j]; j--; }
        wr[j + 1] = x;
    }
    int n_acc = 0, n_b = 0, nvr = 0, nst = 0;
    double max_velocity = (na[A_MAXVEL] > 0) ? na[A_MAXVEL] : V;
    ma[n_acc++] = (na[A_MAXACC] > 0) ? na[A_MAXACC] : A0;
    bi[0] = 0; bv[0] = 0; n_b = 1;
    vi[nvr] = 0; vv[nvr] = max_velocity; nvr++;
    // action point k fires at its smallest candidate sample after the previous action point's sample
    int node_num = 0, wptr = 0, action_idx = 0, last_fire = -1;
    bool actions_dead = false;
    while (true) {
        // next action-point sample (if any)
        int ai = 2147483647;
        if (!actions_dead && action_idx < A) {
            int nc = ev_napc[(size_t)b * A_max + action_idx];
            if (nc > EV_AP_CAND) { status[b] = ST_CAPACITY; return; }
            const int* c = ev_apc + ((size_t)b * A_max + action_idx) * EV_AP_CAND;
            for (int k = 0; k < nc; k++) if (c[k] > last_fire && c[k] < ai) ai = c[k];
            if (ai == 2147483647) actions_dead = true;      // this action point never fires -> none after it does
        }
        int wi = (wptr < nw) ? wr[wptr] : 2147483647;
        if (wi == 2147483647 && ai == 2147483647) break;
        int i = wi < ai ? wi : ai;
        bool stop = false;
        if (wi == i) {                       // node crossing first (:124-140)
            wptr++;
            node_num += 1;
            const double* a = na + (size_t)node_num * NA;
            if (nf[node_num] & F_STOP) stop = true;
            max_velocity = (a[A_MAXVEL] > 0) ? a[A_MAXVEL] : V;
            ma[n_acc++] = (a[A_MAXACC] > 0) ? a[A_MAXACC] : A0;
            if (node_num < n - 1) {
                if (bi[n_b - 1] == i) bv[n_b - 1] = n_acc - 1;
                else { bi[n_b] = i; bv[n_b] = n_acc - 1; n_b++; }
            }
        }
        if (ai == i) {                       // then the action point (:142-163)
            const double* p = apa + (size_t)action_idx * APA;
            max_velocity = (p[P_MAXVEL] > 0) ? p[P_MAXVEL] : V;
            if (apf[action_idx] & F_STOP) stop = true;
            ma[n_acc++] = (p[P_MAXACC] > 0) ? p[P_MAXACC] : A0;
            if (bi[n_b - 1] == i) bv[n_b - 1] = n_acc - 1;
            else { bi[n_b] = i; bv[n_b] = n_acc - 1; n_b++; }
            action_idx += 1;
            last_fire = i;
        }
        if (stop) si[nst++] = i;
        if (vi[nvr - 1] == i + 1) vv[nvr - 1] = max_velocity;
        else { vi[nvr] = i + 1; vv[nvr] = max_velocity; nvr++; }
    }
    ma[n_acc++] = A0;
    n_ev[2 * b] = n_acc; n_ev[2 * b + 1] = n_b;
    n_vr[2 * b] = nvr; n_vr[2 * b + 1] = nst;
    if (ins_est) {
        // rows the time-domain stage inserts for waits and turn profiles (upper bound: assume every one fires)
        const double w = cons[b * 6 + 5];
        double extra = 0.0;
        for (int i = 0; i < n; i++) {
            const double* a = na + (size_t)i * NA;
            if (a[A_WAIT] > 0) extra += floor(a[A_WAIT] / dt);
            if (a[A_TURN] != 0) {
                double angle = a[A_TURN] * (VAP_PI / 180.0);
                Trapezoid tz = trapezoid_setup(V, A0, fabs(angle) * w / 2, dt);
                extra += (double)tz.K;
            }
        }
        for (int i = 0; i < A; i++) if (apa[i * APA + P_WAIT] > 0) extra += floor(apa[i * APA + P_WAIT] / dt);
        ins_est[b] = (float)extra;
    }
}

// ---- pre-pass (parallel): one 32-byte record per sample and direction -----------------------------------------
// forward  record F[i] = { |kappa_i|, 2|theta_{i+1}-theta_i| (NaN when straight), a_static_i, C_i }
//    a_static = min(max_ang_acc/|k|, 2 acc/(w|k|+2), acc)  (acc when straight);  C_i = min(v0[i+1], vlim_i, cap_i)
// backward record R[i] = { |kappa_i|, 2|theta_{i-1}-theta_i| (NaN when straight), d_static_i, G_i }
//    d_static = min(max_ang_acc/|k|, 2 dec/(w|k|+2), dec);  G_i = min(vlim_i, cap_i)
// A NaN denominator makes the wheel term NaN, which Python's min() ignores -- exactly the straight branch.
__global__ void __launch_bounds__(256) k_prepass(
    const int* __restrict__ status, const double* __restrict__ cons, double end_vel, long long D_cap,
    const int* __restrict__ n_samples, const double* __restrict__ kap, const double* __restrict__ th, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const int* __restrict__ vr_idx, const double* __restrict__ vr_val,
    const int* __restrict__ st_idx, const int* __restrict__ n_vr, double4* __restrict__ recF,
    double4* __restrict__ recR)
{
    extern __shared__ unsigned char s_raw[];
    long long b = blockIdx.y;
    if (status[b] != ST_OK) return;
    long long D = n_samples[b];
    long long i0 = (long long)blockIdx.x * blockDim.x;
    if (i0 >= D) return;
    double* s_ma = reinterpret_cast<double*>(s_raw);
    double* s_vv = s_ma + E_cap;
    int* s_bi = reinterpret_cast<int*>(s_vv + E_cap);
    int* s_bv = s_bi + E_cap;
    int* s_vi = s_bv + E_cap;
    int* s_si = s_vi + E_cap;
    const int n_acc = n_ev[2 * b], n_b = n_ev[2 * b + 1], nvr = n_vr[2 * b], nst = n_vr[2 * b + 1];
    for (int k = threadIdx.x; k < E_cap; k += blockDim.x) {
        s_ma[k] = (k < n_acc) ? max_accels[(size_t)b * E_cap + k] : 0.0;
        s_vv[k] = (k < nvr) ? vr_val[(size_t)b * E_cap + k] : 0.0;
        s_bi[k] = (k < n_b) ? bidx[(size_t)b * E_cap + k] : 2147483647;
        s_bv[k] = (k < n_b) ? bval[(size_t)b * E_cap + k] : 0;
        s_vi[k] = (k < nvr) ? vr_idx[(size_t)b * E_cap + k] : 2147483647;
        s_si[k] = (k < nst) ? st_idx[(size_t)b * E_cap + k] : -1;
    }
    // per-path constants and the block's starting positions in the (sorted) event tables: once per CTA
    __shared__ double s_c[4];
    __shared__ int s_j0[2];
    if (threadIdx.x == 0) {
        const double V_ = cons[b * 6 + 0], A0_ = cons[b * 6 + 1], w_ = cons[b * 6 + 5];
        s_c[0] = 2 * V_ / w_;                 // max_angular_vel   (:81)
        s_c[1] = 2 * A0_ / w_;                // max_angular_accel (:82)
        int jf = 0, jv = 0;
        for (int j = 1; j < n_b; j++) if (s_bi[j] <= (int)i0) jf = j;
        for (int j = 1; j < nvr; j++) if (s_vi[j] <= (int)i0 + 1) jv = j;
        s_j0[0] = jf; s_j0[1] = jv;
    }
    __syncthreads();
    long long i = i0 + threadIdx.x;
    if (i >= D) return;
    const double V = cons[b * 6 + 0], w = cons[b * 6 + 5];
    const double max_angular_vel = s_c[0];
    const double max_angular_accel = s_c[1];
    const size_t row = (size_t)b * D_cap;
    double k = kap[row + i], ak = fabs(k);
    double th_i = th[row + i];
    bool straight = ak < 1e-6;
    double vlim, cap;
    if (straight) vlim = V;
    else {
        double v_ang = max_angular_vel / ak;
        double v_kin = 2 * V / (w * ak + 2);
        // Constraints.max_speed_at_curvature (:23-33) with 2*V/w already evaluated
        double m = (max_angular_vel * V) / (ak * V + max_angular_vel);
        double v_curve = pymin(m, V);
        vlim = pymin(pymin(v_ang, v_kin), v_curve);
    }
    cap = fabs(V / (1 + (w * ak / 2)));
    double G = pymin(vlim, cap);
    // forward regime at step i: last boundary with bidx <= i
    int jf = s_j0[0];
    while (jf + 1 < n_b && s_bi[jf + 1] <= (int)i) jf++;
    double acc_f = s_ma[s_bv[jf]];
    double dec_b = s_ma[s_bv[n_b - 1]];     // the backward pass keeps the forward pass's last max_dec
    if (i < D - 1) {
        double v0n;
        if (i + 1 == D - 1) v0n = end_vel;
        else {
            int jv = s_j0[1];
            while (jv + 1 < nvr && s_vi[jv + 1] <= (int)(i + 1)) jv++;
            v0n = s_vv[jv];
            for (int j = 0; j < nst; j++) if (s_si[j] == (int)(i + 1)) v0n = 0.01;
        }
        double astat, h2;
        if (straight) { astat = acc_f; h2 = __longlong_as_double(0x7ff8000000000000LL); }
        else {
            double a_ang = max_angular_accel / ak;
            double a_kin = 2 * acc_f / (w * ak + 2);
            astat = pymin(pymin(a_ang, a_kin), acc_f);
            h2 = 2 * fabs(th[row + i + 1] - th_i);
        }
        recF[row + i] = make_double4(ak, h2, astat, pymin(v0n, G));
    }
    if (i >= 1) {
        double dstat, h2;
        if (straight) { dstat = dec_b; h2 = __longlong_as_double(0x7ff8000000000000LL); }
        else {
            double d_ang = max_angular_accel / ak;
            double d_kin = 2 * dec_b / (w * ak + 2);
            dstat = pymin(pymin(d_ang, d_kin), dec_b);
            h2 = 2 * fabs(th[row + i - 1] - th_i);
        }
        recR[row + i] = make_double4(ak, h2, dstat, G);
    }
}

// ---- chunk-speculative forward / backward kernels: one CTA per path, one chunk per thread ------------------------
__device__ __forceinline__ bool same_bits(double a, double b)
{
    return __double_as_longlong(a) == __double_as_longlong(b);
}

// (w_i^2 - w_{i-1}^2) / (2|dtheta|) with IEEE results for every special case, but without sending the warp through the
// slow path of the inlined division when the numerator is 0 (cruise), or the denominator is 0 (two samples snapped
// to the same table entry) or NaN (the "straight" marker of the pre-pass).
__device__ __forceinline__ double accel_ang_div(double num, double h2)
{
    // ~44 % of consecutive distance samples snap to the same table entry (h2 == 0), so the special cases are the
    // common ones and must not touch the division at all.
    bool special = !(h2 > 0.0) || num == 0.0 || !(fabs(num) < 1e300);
    double n = special ? 1.0 : num, d = special ? 1.0 : h2;
    asm volatile("" : "+d"(n), "+d"(d));   // keep nvcc from folding the selects back into the division's operands
    double q = n / d;
    if (!special) return q;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (h2 > 0.0) {
        if (num == 0.0) return num;                      // +-0 / positive
        return (num != num) ? qnan : copysign(inf, num); // |num| >= 1e300 or inf: overflow / inf / NaN, as IEEE '/' gives
    }
    if (h2 == 0.0) {                                     // x / +0
        if (num > 0.0) return inf;
        if (num < 0.0) return -inf;
        return qnan;                                     // 0/0 or NaN/0
    }
    return qnan;                                         // NaN denominator (the pre-pass's "straight" marker)
}

// forward step i -> i+1 (motion_profile_generator.py:193-249 with the hoisted terms)
__device__ __forceinline__ double fwd_step(const double4 r, double v, double& wp, double acc, double w, double dd)
{
    double ang_vel = v * r.x;
    double accel_ang = accel_ang_div(ang_vel * ang_vel - wp * wp, r.y);
    double a_wheel = wheel_accel(acc, fabs(accel_ang), w);
    if (a_wheel < 0) a_wheel = 0;
    double a = pymin(r.z, a_wheel);
    double s = sqrt(v * v + 2 * a * dd);
    wp = ang_vel;
    return pymin(r.w, s);
}
// backward step i -> i-1 (:255-311)
__device__ __forceinline__ double bwd_step(const double4 r, double v, double& wp, double acc, double w, double dd,
                                           double vfwd_prev)
{
    double ang_vel = v * r.x;
    double accel_ang = accel_ang_div(ang_vel * ang_vel - wp * wp, r.y);
    double a_wheel = wheel_accel(acc, accel_ang, w);
    if (a_wheel < 0) a_wheel = 0;
    double dcl = pymin(r.z, a_wheel);
    double pv = sqrt(v * v + 2 * dcl * dd);
    wp = ang_vel;
    return pymin(pymin(pv, vfwd_prev), r.w);
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk-speculative passes, CTA = PB paths x NT chunks (thread (p, c) = path p of the CTA, chunk c).
//
// Sweep 1: every thread runs its own chunk from the guessed state (lockstep, all lanes busy).
// Fix-up rounds: only a few chunks per path have to be re-run (their predecessor's end state differs bitwise from the
// state they started from).  Those (path, chunk) items are pushed into a shared-memory queue and re-run by the FIRST
// threads of the CTA, so that the re-runs of all PB paths share a few dense warps instead of every path keeping a
// mostly idle warp spinning.  A re-run stops as soon as two consecutive velocities equal the stored ones bitwise.
// Rounds end when the queue stays empty; then, by induction from the first chunk, every value is the serial value.
// ------------------------------------------------------------------------------------------------------------------
struct ChunkCtx {                 // per path; built in registers from the kernel parameters
    long long D, Lc;
    int nch, n_b, ok;
    double w;
    const double4* rec;
    const double* vin;            // backward: forward velocities
    double* vout;
    const double* ma;
    const int* bi;
    const int* bv;
};

// Forward pass.  Steps i = 0 .. D-2; chunk c owns steps [c*Lc, min((c+1)*Lc, D-1)); step i writes vel_f[i+1].
__global__ void __launch_bounds__(256) k_fwd_chunked(
    long long B, int NT, const int* __restrict__ status, const double* __restrict__ cons, double dd, double start_vel,
    long long D_cap, const int* __restrict__ n_samples, const double4* __restrict__ recF, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, double* __restrict__ vel_f, int* __restrict__ rounds_out, int warm)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int NTH = blockDim.x, PB = NTH / NT;
    const int tid = threadIdx.x, p = tid / NT, c = tid - p * NT;
    double* s_endv = reinterpret_cast<double*>(s_mem);           // [NTH] end state of every chunk
    double* s_endw = s_endv + NTH;
    double* s_usev = s_endw + NTH;                                // [NTH] start state every chunk last used
    double* s_usew = s_usev + NTH;
    int* s_queue = reinterpret_cast<int*>(s_usew + NTH);          // [NTH]
    __shared__ int s_qn;
    // path context from the kernel parameters (pointers stay provably global: ld.global.nc, no generic loads)
    auto make_ctx = [&](int pp) -> ChunkCtx {
        ChunkCtx x;
        const long long bb = (long long)blockIdx.x * PB + pp;
        x.ok = (bb < B) && status[bb] == ST_OK;
        x.D = x.ok ? n_samples[bb] : 1;
        long long steps = x.D - 1;
        x.Lc = steps > 0 ? (steps + NT - 1) / NT : 1;
        x.nch = (x.ok && steps > 0) ? (int)((steps + x.Lc - 1) / x.Lc) : 0;
        const long long bs = x.ok ? bb : 0;
        x.n_b = x.ok ? n_ev[2 * bs + 1] : 0;
        x.w = cons[bs * 6 + 5];
        x.rec = recF + (size_t)bs * D_cap;
        x.vin = nullptr;
        x.vout = vel_f + (size_t)bs * D_cap;
        x.ma = max_accels + (size_t)bs * E_cap;
        x.bi = bidx + (size_t)bs * E_cap;
        x.bv = bval + (size_t)bs * E_cap;
        return x;
    };
    const long long b = (long long)blockIdx.x * PB + p;
    const ChunkCtx mine = make_ctx(p);
    if (c == 0 && mine.ok) mine.vout[0] = start_vel;
    if (tid == 0) s_qn = 0;
    __syncthreads();

    // runs chunk cc of path context x from state (v, wp); mode 0: first sweep (plain stores), mode 1: re-run with
    // bitwise merge detection against the stored velocities.  Returns true when the re-run merged (end state unchanged).
    auto run_chunk = [&](const ChunkCtx& x, int cc, double& v, double& wp, bool rerun, bool prev_same) -> bool {
        const long long steps = x.D - 1;
        const long long lo = (long long)cc * x.Lc;
        const long long hi = (lo + x.Lc < steps) ? lo + x.Lc : steps;
        const double4* F = x.rec;
        double* vf = x.vout;
        int j = 1;
        double acc = __ldg(x.ma + x.bv[0]);
        while (j < x.n_b && x.bi[j] <= (int)lo) { acc = __ldg(x.ma + x.bv[j]); j++; }
        int nb_next = (j < x.n_b) ? x.bi[j] : 2147483647;
        double acc_next = (j < x.n_b) ? __ldg(x.ma + x.bv[j]) : acc;   // fetched one regime ahead: no load on the per-step path
        double4 r = ldg_d4(F + lo);
        double4 r1 = (lo + 1 < hi) ? ldg_d4(F + lo + 1) : r;
        if (!rerun) {
            for (long long i = lo; i < hi; i++) {
                double4 r2 = (i + 2 < hi) ? ldg_d4(F + i + 2) : r1;          // records run two steps ahead of their use
                if ((int)i == nb_next) {
                    acc = acc_next; j++;
                    nb_next = (j < x.n_b) ? x.bi[j] : 2147483647;
                    acc_next = (j < x.n_b) ? __ldg(x.ma + x.bv[j]) : acc;
                }
                v = fwd_step(r, v, wp, acc, x.w, dd);
                vf[i + 1] = v;
                r = r1; r1 = r2;
            }
            return false;
        }
        double old = vf[lo + 1];
        double old1 = (lo + 1 < hi) ? vf[lo + 2] : 0.0;
        for (long long i = lo; i < hi; i++) {
            double4 r2 = (i + 2 < hi) ? ldg_d4(F + i + 2) : r1;
            double old2 = (i + 2 < hi) ? vf[i + 3] : 0.0;
            if ((int)i == nb_next) {
                acc = acc_next; j++;
                nb_next = (j < x.n_b) ? x.bi[j] : 2147483647;
                acc_next = (j < x.n_b) ? __ldg(x.ma + x.bv[j]) : acc;
            }
            v = fwd_step(r, v, wp, acc, x.w, dd);
            bool same = same_bits(old, v);
            if (same && prev_same) return true;        // state (v[i+1], v[i]*|k_i|) equals the old run's: rest is unchanged
            vf[i + 1] = v;
            prev_same = same;
            r = r1; r1 = r2; old = old1; old1 = old2;
        }
        return false;
    };

    // ---- sweep 1.  Chunk c > 0 starts `warm` steps BEFORE its own range from the guess "the state-independent caps bind
    // on the two samples before that point" and only computes (no stores) until it reaches its range: a wrong guess
    // survives for at most one acceleration ramp, so after the warm-up the state is usually already the true one and the
    // fix-up round has nothing to re-run.  (Exactness never depends on the guess: the rounds below verify bitwise.)
    {
        const ChunkCtx& x = mine;
        const bool active = x.ok && c < x.nch;
        double v = start_vel, wp = 0.0;
        if (active) {
            const long long lo = (long long)c * x.Lc;
            if (c > 0) {
                const long long wl = (warm < lo - 2) ? warm : ((lo - 2 > 0) ? lo - 2 : 0);
                const long long s0 = lo - wl;                       // first warm-up step (>= 2 unless lo < 2)
                double vm1 = (s0 >= 2) ? x.rec[s0 - 2].w : start_vel;
                double4 fm1 = x.rec[s0 - 1];
                v = fm1.w;
                wp = vm1 * fm1.x;
                if (wl > 0) {
                    int j = 1;
                    double acc = x.ma[x.bv[0]];
                    while (j < x.n_b && x.bi[j] <= (int)s0) { acc = x.ma[x.bv[j]]; j++; }
                    int nb_next = (j < x.n_b) ? x.bi[j] : 2147483647;
                    double4 r = x.rec[s0];
                    double4 r1 = (s0 + 1 < lo) ? x.rec[s0 + 1] : r;
                    for (long long i = s0; i < lo; i++) {
                        double4 r2 = (i + 2 < lo) ? x.rec[i + 2] : r1;
                        if ((int)i == nb_next) { acc = x.ma[x.bv[j]]; j++; nb_next = (j < x.n_b) ? x.bi[j] : 2147483647; }
                        v = fwd_step(r, v, wp, acc, x.w, dd);
                        r = r1; r1 = r2;
                    }
                }
            }
            s_usev[tid] = v; s_usew[tid] = wp;
            run_chunk(x, c, v, wp, false, false);
        }
        s_endv[tid] = v; s_endw[tid] = wp;
    }
    __syncthreads();

    // ---- fix-up rounds
    int rounds = 0;
    for (int round = 1; round < NT; round++) {
        {
            const ChunkCtx& x = mine;
            if (x.ok && c < x.nch && c >= round) {
                if (!(same_bits(s_endv[tid - 1], s_usev[tid]) && same_bits(s_endw[tid - 1], s_usew[tid])))
                    s_queue[atomicAdd(&s_qn, 1)] = tid;
            }
        }
        __syncthreads();
        const int n = s_qn;
        if (n == 0) break;
        rounds = round;
        // read the inputs of this round before anybody publishes new end states
        int item = -1;
        double in_v = 0.0, in_w = 0.0, old_usev = 0.0;
        if (tid < n) { item = s_queue[tid]; in_v = s_endv[item - 1]; in_w = s_endw[item - 1]; old_usev = s_usev[item]; }
        __syncthreads();
        if (tid == 0) s_qn = 0;
        if (item >= 0) {
            const int pp = item / NT, cc = item - pp * NT;
            const ChunkCtx x = (pp == p) ? mine : make_ctx(pp);
            double v = in_v, wp = in_w;
            bool merged = run_chunk(x, cc, v, wp, true, same_bits(in_v, old_usev));
            s_usev[item] = in_v; s_usew[item] = in_w;
            if (!merged) { s_endv[item] = v; s_endw[item] = wp; }
        }
        __syncthreads();
    }
    if (c == 0 && rounds_out && mine.ok) rounds_out[2 * b] = rounds;
}

// Backward pass.  Steps i = D-1 .. 1 (step i writes vel[i-1]); chunk c owns the c-th block of steps counted from the
// end.  Reads vel_f (forward result) and writes vel (final); vel[D-1] = end_vel.  Also accumulates the travel-time
// estimate used to size the time-domain outputs.
__global__ void __launch_bounds__(256) k_bwd_chunked(
    long long B, int NT, const int* __restrict__ status, const double* __restrict__ cons, double dd, double dt,
    double end_vel, long long D_cap, const int* __restrict__ n_samples, const double4* __restrict__ recR, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const double* __restrict__ vel_f, double* __restrict__ vel,
    float* __restrict__ t_est, int* __restrict__ rounds_out, int warm)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int NTH = blockDim.x, PB = NTH / NT;
    const int tid = threadIdx.x, p = tid / NT, c = tid - p * NT;
    double* s_endv = reinterpret_cast<double*>(s_mem);
    double* s_endw = s_endv + NTH;
    double* s_usev = s_endw + NTH;
    double* s_usew = s_usev + NTH;
    int* s_queue = reinterpret_cast<int*>(s_usew + NTH);
    __shared__ int s_qn;
    auto make_ctx = [&](int pp) -> ChunkCtx {
        ChunkCtx x;
        const long long bb = (long long)blockIdx.x * PB + pp;
        x.ok = (bb < B) && status[bb] == ST_OK;
        x.D = x.ok ? n_samples[bb] : 1;
        long long steps = x.D - 1;
        x.Lc = steps > 0 ? (steps + NT - 1) / NT : 1;
        x.nch = (x.ok && steps > 0) ? (int)((steps + x.Lc - 1) / x.Lc) : 0;
        const long long bs = x.ok ? bb : 0;
        x.n_b = x.ok ? n_ev[2 * bs + 1] : 0;
        x.w = cons[bs * 6 + 5];
        x.rec = recR + (size_t)bs * D_cap;
        x.vin = vel_f + (size_t)bs * D_cap;
        x.vout = vel + (size_t)bs * D_cap;
        x.ma = max_accels + (size_t)bs * E_cap;
        x.bi = bidx + (size_t)bs * E_cap;
        x.bv = bval + (size_t)bs * E_cap;
        return x;
    };
    const long long b = (long long)blockIdx.x * PB + p;
    const ChunkCtx mine = make_ctx(p);
    if (c == 0 && mine.ok) mine.vout[mine.D - 1] = end_vel;
    if (tid == 0) s_qn = 0;
    __syncthreads();

    // chunk cc: steps i = hi, hi-1, ..., lo+1 with hi = D-1 - cc*Lc, lo = max(hi - Lc, 0)
    auto run_chunk = [&](const ChunkCtx& x, int cc, double& v, double& wp, bool rerun, bool prev_same) -> bool {
        const long long hi = (x.D - 1) - (long long)cc * x.Lc;
        const long long lo = (hi - x.Lc > 0) ? hi - x.Lc : 0;
        const double4* R = x.rec;
        const double* vf = x.vin;
        double* vo = x.vout;
        // regime at chunk start (walking down from D-1): the smallest boundary index > hi was the last one applied
        double acc = __ldg(x.ma + x.bv[x.n_b - 1]);
        int j = x.n_b - 1;
        while (j >= 0 && x.bi[j] > (int)hi) { acc = __ldg(x.ma + x.bv[j] + 1); j--; }
        int nb_next = (j >= 0) ? x.bi[j] : -1;
        double acc_next = (j >= 0) ? __ldg(x.ma + x.bv[j] + 1) : acc;   // fetched one regime ahead
        double4 r = ldg_d4(R + hi);
        double vfp = __ldg(vf + hi - 1);
        double4 r1 = (hi - 1 > lo) ? ldg_d4(R + hi - 1) : r;
        double vf1 = (hi - 1 > lo) ? __ldg(vf + hi - 2) : 0.0;
        if (!rerun) {
            for (long long i = hi; i > lo; i--) {
                double4 r2 = (i - 2 > lo) ? ldg_d4(R + i - 2) : r1;          // loads run two steps ahead of their use
                double vf2 = (i - 2 > lo) ? __ldg(vf + i - 3) : 0.0;
                if ((int)i == nb_next) {
                    acc = acc_next; j--;
                    nb_next = (j >= 0) ? x.bi[j] : -1;
                    acc_next = (j >= 0) ? __ldg(x.ma + x.bv[j] + 1) : acc;
                }
                v = bwd_step(r, v, wp, acc, x.w, dd, vfp);
                vo[i - 1] = v;
                r = r1; vfp = vf1; r1 = r2; vf1 = vf2;
            }
            return false;
        }
        double old = vo[hi - 1];
        double old1 = (hi - 1 > lo) ? vo[hi - 2] : 0.0;
        for (long long i = hi; i > lo; i--) {
            double4 r2 = (i - 2 > lo) ? ldg_d4(R + i - 2) : r1;
            double vf2 = (i - 2 > lo) ? __ldg(vf + i - 3) : 0.0;
            double old2 = (i - 2 > lo) ? vo[i - 3] : 0.0;
            if ((int)i == nb_next) {
                acc = acc_next; j--;
                nb_next = (j >= 0) ? x.bi[j] : -1;
                acc_next = (j >= 0) ? __ldg(x.ma + x.bv[j] + 1) : acc;
            }
            v = bwd_step(r, v, wp, acc, x.w, dd, vfp);
            bool same = same_bits(old, v);
            if (same && prev_same) return true;
            vo[i - 1] = v;
            prev_same = same;
            r = r1; vfp = vf1; old = old1; r1 = r2; vf1 = vf2; old1 = old2;
        }
        return false;
    };

    // ---- sweep 1
    {
        const ChunkCtx& x = mine;
        const bool active = x.ok && c < x.nch;
        double v = end_vel, wp = 0.0;
        if (active) {
            const long long hi = (x.D - 1) - (long long)c * x.Lc;
            if (c > 0) {
                // warm-up (see the forward kernel): start `warm` steps above the chunk from the guess
                // v[s0] = min(vel_f[s0], G[s0+1]) (the state-independent part of what step s0+1 produces)
                const long long room = (x.D - 1) - hi - 2;
                const long long wl = (warm < room) ? warm : (room > 0 ? room : 0);
                const long long s0 = hi + wl;
                double4 r1 = x.rec[s0 + 1];
                v = pymin(x.vin[s0], r1.w);
                double vp1 = (s0 + 2 <= x.D - 1) ? pymin(x.vin[s0 + 1], x.rec[s0 + 2].w) : end_vel;
                wp = vp1 * r1.x;
                if (wl > 0) {
                    double acc = x.ma[x.bv[x.n_b - 1]];
                    int j = x.n_b - 1;
                    while (j >= 0 && x.bi[j] > (int)s0) { acc = x.ma[x.bv[j] + 1]; j--; }
                    int nb_next = (j >= 0) ? x.bi[j] : -1;
                    double4 r = x.rec[s0];
                    double vfp = x.vin[s0 - 1];
                    double4 ra = (s0 - 1 > hi) ? x.rec[s0 - 1] : r;
                    double vfa = (s0 - 1 > hi) ? x.vin[s0 - 2] : 0.0;
                    for (long long i = s0; i > hi; i--) {
                        double4 rb = (i - 2 > hi) ? x.rec[i - 2] : ra;
                        double vfb = (i - 2 > hi) ? x.vin[i - 3] : 0.0;
                        if ((int)i == nb_next) { acc = x.ma[x.bv[j] + 1]; j--; nb_next = (j >= 0) ? x.bi[j] : -1; }
                        v = bwd_step(r, v, wp, acc, x.w, dd, vfp);
                        r = ra; vfp = vfa; ra = rb; vfa = vfb;
                    }
                }
            }
            s_usev[tid] = v; s_usew[tid] = wp;
            run_chunk(x, c, v, wp, false, false);
        }
        s_endv[tid] = v; s_endw[tid] = wp;
    }
    __syncthreads();

    // ---- fix-up rounds (states flow from chunk c-1 to chunk c, as in the forward kernel)
    int rounds = 0;
    for (int round = 1; round < NT; round++) {
        {
            const ChunkCtx& x = mine;
            if (x.ok && c < x.nch && c >= round) {
                if (!(same_bits(s_endv[tid - 1], s_usev[tid]) && same_bits(s_endw[tid - 1], s_usew[tid])))
                    s_queue[atomicAdd(&s_qn, 1)] = tid;
            }
        }
        __syncthreads();
        const int n = s_qn;
        if (n == 0) break;
        rounds = round;
        int item = -1;
        double in_v = 0.0, in_w = 0.0, old_usev = 0.0;
        if (tid < n) { item = s_queue[tid]; in_v = s_endv[item - 1]; in_w = s_endw[item - 1]; old_usev = s_usev[item]; }
        __syncthreads();
        if (tid == 0) s_qn = 0;
        if (item >= 0) {
            const int pp = item / NT, cc = item - pp * NT;
            const ChunkCtx x = (pp == p) ? mine : make_ctx(pp);
            double v = in_v, wp = in_w;
            bool merged = run_chunk(x, cc, v, wp, true, same_bits(in_v, old_usev));
            s_usev[item] = in_v; s_usew[item] = in_w;
            if (!merged) { s_endv[item] = v; s_endw[item] = wp; }
        }
        __syncthreads();
    }
    if (c == 0 && rounds_out && mine.ok) rounds_out[2 * b + 1] = rounds;

    // ---- travel-time estimate (single precision is plenty: it only sizes buffers)
    __syncthreads();
    float est = 0.f;
    {
        const ChunkCtx& x = mine;
        if (x.ok && c < x.nch) {
            const long long hi = (x.D - 1) - (long long)c * x.Lc;
            const long long lo = (hi - x.Lc > 0) ? hi - x.Lc : 0;
            for (long long i = hi; i > lo; i--) {
                float vm = 0.5f * ((float)x.vout[i] + (float)x.vout[i - 1]);
                est += __fdividef((float)dd, fmaxf(vm, 0.05f) * (float)dt);
            }
        }
    }
    float* s_f = reinterpret_cast<float*>(s_endv);      // end states are no longer needed
    __syncthreads();
    s_f[tid] = est;
    __syncthreads();
    if (c == 0 && t_est && b < B) {
        float tot = 0.f;
        if (mine.ok) for (int k = 0; k < NT; k++) tot += s_f[p * NT + k];
        t_est[b] = tot;
    }
}
