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
    int* __restrict__ ev_wrap, int* __restrict__ ev_nwrap, int* __restrict__ ev_apc, int* __restrict__ ev_napc,
    const int* __restrict__ lut_inv)
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
    const int* inv = lut_inv ? lut_inv + (size_t)b * (Q_cap + LUT_INV_HDR + 2) : nullptr;
    auto d2t = [&](double d) { return inv ? distance_to_time_inv(ld, lt, inv, Q, L, n, d) : distance_to_time32(ld, lt, Q, L, n, d); };
    double t = 0.0;
    if (i < D) {
        t = (i == D - 1) ? (double)(n - 1) : d2t(dgrid[i]);
        double k, h;
        snap_gather2_32(prop_k + (size_t)b * P_cap, prop_h + (size_t)b * P_cap, t, pg, k, h);
        size_t o = (size_t)b * D_cap + i;
        t_out[o] = t; kap[o] = k; th[o] = h;
    }
    s_t[threadIdx.x + 1] = t;
    if (threadIdx.x == 0) s_t[0] = (i0 > 0) ? d2t(dgrid[i0 - 1]) : 0.0;   // prev_t of sample 0 is 0
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

__device__ __forceinline__ double recip_for_pass(double g);   // defined with the pass kernels below

// ---- chunk-interleaved layout of everything the passes stream ---------------------------------------------------------
// A path's D-1 steps ("edges" e = 0 .. D-2: forward step e goes from sample e to e+1, backward step e+1 from sample e+1
// to e) are cut into NT chunks of Lc = ceil((D-1)/NT) edges; chunk c owns edges [c*Lc, min((c+1)*Lc, D-1)).  Thread c of a
// pass CTA walks chunk c, so at any moment the warp needs edge s of 32 different chunks: the per-edge data is stored at
//      slot(e) = (e % Lc) * NT + e / Lc          (row = position inside the chunk, column = chunk)
// which makes every warp-wide load of the passes one contiguous, fully used run of memory (1 KB of records, 256 B of
// reciprocals / forward velocities).  Rows of RS = D_cap + 256 slots per path; slot RS-1 holds the forward velocity of
// the last sample.
__device__ __forceinline__ int chunk_len(int steps, int NT) { return (steps + NT - 1) / NT; }
__device__ __forceinline__ int edge_slot(int e, int Lc, int NT) { const int c = e / Lc; return (e - c * Lc) * NT + c; }

// ---- pre-pass (parallel): one 32-byte record per edge and direction ----------------------------------------------------
// forward  record of edge e (sample i = e):     { |kappa_i|, 2|theta_{i+1}-theta_i| (NaN when straight), a_static_i, C_i }
//    a_static = min(max_ang_acc/|k|, 2 acc/(w|k|+2), acc)  (acc when straight);  C_i = min(v0[i+1], vlim_i, cap_i)
// backward record of edge e (sample i = e + 1): { |kappa_i|, 2|theta_{i-1}-theta_i| (NaN when straight), d_static_i, G_i }
//    d_static = min(max_ang_acc/|k|, 2 dec/(w|k|+2), dec);  G_i = min(vlim_i, cap_i)
// rg of edge e: recip_for_pass(2|theta_{e+1}-theta_e|), shared by both directions.
// A NaN denominator makes the wheel term NaN, which Python's min() ignores -- exactly the straight branch.
// One thread per SLOT (coalesced stores; the 8-byte gathers of kappa / theta are 8 rows deep per CTA, so every fetched
// sector is used by the CTA's other warps).  The thread of sample i writes the forward record of edge i and the backward
// record of edge i-1; the thread of the last edge also makes the backward record of the final sample.
struct PrepassTables {
    const double* ma; const double* vv; const int* bi; const int* bv; const int* vi; const int* si;
    int n_b, nvr, nst;
};
__device__ __forceinline__ int last_le(const int* a, int n, int key)      // last j in [0, n) with a[j] <= key (a[0] <= key)
{
    int j = 0;
    if (n <= 8) { while (j + 1 < n && a[j + 1] <= key) j++; return j; }
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (a[mid] <= key) lo = mid; else hi = mid - 1; }
    return lo;
}
// a / b for several numerators over one denominator: y = RN(1/b) once, then per quotient one multiplication and two fused
// residual corrections (Markstein: with the correctly rounded reciprocal the second correction gives the correctly rounded
// quotient unless b's significand is all ones).  Operands outside the proven range take the ordinary division.
struct SharedRecip { double b, y; bool ok; };
__device__ __forceinline__ SharedRecip shared_recip(double b)
{
    SharedRecip r;
    const long long bits = __double_as_longlong(b);
    r.ok = (b > 1e-150) && (b < 1e150) && ((bits & 0x000FFFFFFFFFFFFFLL) != 0x000FFFFFFFFFFFFFLL);
    r.b = b;
    r.y = 1.0 / b;
    return r;
}
__device__ __forceinline__ double div_shared(double a, const SharedRecip& r)
{
    const double aa = fabs(a);
    if (!(r.ok && aa > 1e-150 && aa < 1e150)) return a / r.b;
    double q = a * r.y;
    q = fma(fma(-r.b, q, a), r.y, q);
    q = fma(fma(-r.b, q, a), r.y, q);
    return q;
}
__device__ __forceinline__ void prepass_sample(const PrepassTables& T, double V, double w, double max_angular_vel,
                                               double max_angular_accel, double end_vel, int D, int i, double k,
                                               double th_i, double th_next, double th_prev, bool wantF, bool wantR,
                                               double4& F, double4& R)
{
    const double ak = fabs(k);
    const bool straight = ak < 1e-6;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // w|k| + 2 = 2 (1 + w|k|/2) bit for bit (halving and doubling are exact), so the wheel-speed cap
    // |V / (1 + w|k|/2)| (:239) IS 2V / (w|k| + 2) (:214): one quotient serves both.
    const SharedRecip rden = shared_recip(w * ak + 2);
    const double v_kin = div_shared(2 * V, rden);
    const double cap = fabs(v_kin);
    double vlim, a_ang = 0.0;
    SharedRecip rak;
    if (straight) vlim = V;
    else {
        rak = shared_recip(ak);
        double v_ang = div_shared(max_angular_vel, rak);
        a_ang = div_shared(max_angular_accel, rak);
        // Constraints.max_speed_at_curvature (:23-33) with 2*V/w already evaluated
        double m = (max_angular_vel * V) / (ak * V + max_angular_vel);
        double v_curve = pymin(m, V);
        vlim = pymin(pymin(v_ang, v_kin), v_curve);
    }
    const double G = pymin(vlim, cap);
    const double dec_b = T.ma[T.bv[T.n_b - 1]];     // the backward pass keeps the forward pass's last max_dec
    double acc_f = dec_b, a_kin = 0.0;
    if (wantF) {
        if (T.n_b > 1) acc_f = T.ma[T.bv[last_le(T.bi, T.n_b, i)]];    // forward regime at step i: last boundary <= i
        double v0n;
        if (i + 1 == D - 1) v0n = end_vel;
        else {
            v0n = (T.nvr > 1) ? T.vv[last_le(T.vi, T.nvr, i + 1)] : T.vv[0];
            for (int j = 0; j < T.nst; j++) if (T.si[j] == i + 1) v0n = 0.01;
        }
        double astat, h2;
        if (straight) { astat = acc_f; h2 = qnan; }
        else {
            a_kin = div_shared(2 * acc_f, rden);
            astat = pymin(pymin(a_ang, a_kin), acc_f);
            h2 = 2 * fabs(th_next - th_i);
        }
        F = make_double4(ak, h2, astat, pymin(v0n, G));
    }
    if (wantR) {
        double dstat, h2;
        if (straight) { dstat = dec_b; h2 = qnan; }
        else {
            double d_kin = (wantF && acc_f == dec_b) ? a_kin : div_shared(2 * dec_b, rden);
            dstat = pymin(pymin(a_ang, d_kin), dec_b);
            h2 = 2 * fabs(th_prev - th_i);
        }
        R = make_double4(ak, h2, dstat, G);
    }
}

__global__ void __launch_bounds__(256) k_prepass(
    const int* __restrict__ status, const double* __restrict__ cons, double end_vel, long long D_cap,
    const int* __restrict__ n_samples, const double* __restrict__ kap, const double* __restrict__ th, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const int* __restrict__ vr_idx, const double* __restrict__ vr_val,
    const int* __restrict__ st_idx, const int* __restrict__ n_vr, int NT, long long RS, double4* __restrict__ recF,
    double4* __restrict__ recR, double* __restrict__ rg)
{
    extern __shared__ unsigned char s_raw[];
    const long long b = blockIdx.y;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int steps = D - 1;
    if (steps <= 0) return;
    const int Lc = chunk_len(steps, NT);
    const int j0 = blockIdx.x * blockDim.x;
    if (j0 >= Lc * NT) return;
    double* s_ma = reinterpret_cast<double*>(s_raw);
    double* s_vv = s_ma + E_cap;
    int* s_bi = reinterpret_cast<int*>(s_vv + E_cap);
    int* s_bv = s_bi + E_cap;
    int* s_vi = s_bv + E_cap;
    int* s_si = s_vi + E_cap;
    PrepassTables T;
    const int n_acc = n_ev[2 * b];
    T.n_b = n_ev[2 * b + 1]; T.nvr = n_vr[2 * b]; T.nst = n_vr[2 * b + 1];
    for (int k = threadIdx.x; k < E_cap; k += blockDim.x) {
        s_ma[k] = (k < n_acc) ? max_accels[(size_t)b * E_cap + k] : 0.0;
        s_vv[k] = (k < T.nvr) ? vr_val[(size_t)b * E_cap + k] : 0.0;
        s_bi[k] = (k < T.n_b) ? bidx[(size_t)b * E_cap + k] : 2147483647;
        s_bv[k] = (k < T.n_b) ? bval[(size_t)b * E_cap + k] : 0;
        s_vi[k] = (k < T.nvr) ? vr_idx[(size_t)b * E_cap + k] : 2147483647;
        s_si[k] = (k < T.nst) ? st_idx[(size_t)b * E_cap + k] : -1;
    }
    T.ma = s_ma; T.vv = s_vv; T.bi = s_bi; T.bv = s_bv; T.vi = s_vi; T.si = s_si;
    // kappa / theta of this CTA's slots: rows s0 .. s0+RW-1 of every column plus one halo row on each side, fetched with the
    // lanes running ALONG a column (contiguous samples) into shared memory
    const double* kr = kap + (size_t)b * D_cap;
    const double* tr = th + (size_t)b * D_cap;
    const int sh = 31 - __clz(NT);                   // NT is a power of two
    const int RW = blockDim.x >> sh;                 // rows per CTA
    const int st = (RW + 2) | 1;                     // odd tile stride
    double* t_th = reinterpret_cast<double*>(s_si + E_cap);            // 32 * E_cap bytes in: 8-byte aligned
    double* t_k = t_th + NT * st;
    const int s0 = j0 >> sh;
    for (int q = threadIdx.x; q < NT * (RW + 2); q += blockDim.x) {
        const int cc = q / (RW + 2), r = q - cc * (RW + 2);
        int ee = cc * Lc + s0 - 1 + r;
        ee = ee < 0 ? 0 : (ee > D - 1 ? D - 1 : ee);
        t_th[cc * st + r] = tr[ee];
        if (r >= 1 && r <= RW) t_k[cc * st + r] = kr[ee];
    }
    __syncthreads();
    const int j = j0 + threadIdx.x;
    const int s = j >> sh, c = j & (NT - 1);
    const int e = c * Lc + s;                  // this slot's edge = the sample whose terms this thread evaluates
    if (s >= Lc || e >= steps) return;
    const double V = cons[b * 6 + 0], A0 = cons[b * 6 + 1], w = cons[b * 6 + 5];
    const double max_angular_vel = 2 * V / w;       // (:81)
    const double max_angular_accel = 2 * A0 / w;    // (:82)
    const size_t row = (size_t)b * RS;
    const int tl = c * st + (s - s0) + 1;
    const double th_i = t_th[tl], th_n = t_th[tl + 1];
    const double th_p = t_th[tl - 1];
    double4 F, R;
    prepass_sample(T, V, w, max_angular_vel, max_angular_accel, end_vel, D, e, t_k[tl], th_i, th_n, th_p, true, e >= 1, F, R);
    recF[row + j] = F;
    rg[row + j] = recip_for_pass(2 * fabs(th_n - th_i));
    if (e >= 1) recR[row + ((s >= 1) ? j - NT : (Lc - 1) * NT + c - 1)] = R;     // slot of edge e-1
    if (e == steps - 1) {                                                         // backward record of the final sample
        prepass_sample(T, V, w, max_angular_vel, max_angular_accel, end_vel, D, D - 1, kr[D - 1], th_n, 0.0, th_i, false, true, F, R);
        recR[row + j] = R;
    }
}

// ---- chunk-speculative forward / backward kernels: one CTA per path, one chunk per thread ------------------------
__device__ __forceinline__ bool same_bits(double a, double b)
{
    return __double_as_longlong(a) == __double_as_longlong(b);
}

// (w_i^2 - w_{i-1}^2) / (2|dtheta|) with IEEE results for every special case, but without sending the warp through the
// slow path of the inlined division when the numerator is 0 (cruise), or the denominator is 0 (two samples snapped
// to the same table entry) or NaN (the "straight" marker of the pre-pass).  Generic path (rare inputs only).
__device__ __noinline__ double accel_ang_div(double num, double h2)
{
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

// Reciprocal of the division's denominator g = 2|dtheta|, made ONCE per sample by the pre-pass (it does not depend on the
// velocity state), so that the state-dependent chain of a pass step holds one multiplication and two fused residual
// corrections instead of a reciprocal refinement:
//    > 0 and finite : RN(1/g), g a normal number whose significand is not all ones (Markstein's exception)
//    +inf           : g == 0 (two samples snapped to the same heading entry; ~44 % of the samples)
//    -1             : anything else (g NaN / subnormal / all-ones significand): the pass takes the generic division
__device__ __forceinline__ double recip_for_pass(double g)
{
    const long long bits = __double_as_longlong(g);
    const bool safe = (g >= 2.2250738585072014e-308) && (g < 1e300) &&
                      ((bits & 0x000FFFFFFFFFFFFFLL) != 0x000FFFFFFFFFFFFFLL);
    double d = safe ? g : 1.0;
    asm volatile("" : "+d"(d));            // the division must never see the special values (warp-wide slow path)
    const double r = 1.0 / d;
    return safe ? r : ((g == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : -1.0);
}

// num / h2 given rce = recip_for_pass(h2) (or NaN when h2 is the straight marker).  q0 = num * r, then two residual
// corrections: after the first the quotient is faithful, and with the correctly rounded reciprocal one more gives the
// correctly rounded quotient (Markstein), i.e. exactly what '/' returns (the sign of a zero quotient is not kept; it
// cannot reach any output).  h2 == 0 and the straight marker fall out of the multiplication: x * inf and x * NaN are
// what x / 0 and x / NaN give.  tests/test_gpu_parity.py::test_recip_division_is_ieee checks 2^30 random pairs.
__device__ __forceinline__ double accel_ang_fast(double num, double h2, double rce)
{
    const double q0 = num * rce;
    double q = fma(fma(-h2, q0, num), rce, q0);
    q = fma(fma(-h2, q, num), rce, q);
    const bool corr = rce <= 1.7976931348623157e308;          // false for NaN and +inf
    const double an = fabs(num);
    const bool in_range = (an < 1e200) && ((an > 1e-200) || (num == 0.0));
    if (rce < 0.0 || (corr && !in_range)) return accel_ang_div(num, h2);     // never taken on sane inputs
    return corr ? q : q0;
}

// forward step i -> i+1 (motion_profile_generator.py:193-249 with the hoisted terms).  sq carries (v_{i-1}|k_{i-1}|)^2.
__device__ __forceinline__ double fwd_step(const double4 r, double rc, double v, double& sq, double acc, double hw, double dd)
{
    const double ang_vel = v * r.x;
    const double sqn = ang_vel * ang_vel;
    const double rce = (r.y == r.y) ? rc : r.y;
    const double accel_ang = accel_ang_fast(sqn - sq, r.y, rce);
    const double x = fabs(accel_ang) * hw;                   // ang * w / 2: halving is exact, so (ang*w)/2 == ang*(w/2)
    const double l = acc + x, rr = acc - x;                  // wheel_accel (:52-59)
    double a_wheel = (fabs(l) < fabs(rr)) ? l : rr;
    if (a_wheel < 0) a_wheel = 0;
    const double a = pymin(r.z, a_wheel);
    const double s = sqrt(v * v + 2 * a * dd);
    sq = sqn;
    return pymin(r.w, s);
}
// backward step i -> i-1 (:255-311); m = min(vel_f[i-1], G_i) (state-independent part of the three-way min)
__device__ __forceinline__ double bwd_step(const double4 r, double rc, double v, double& sq, double acc, double hw, double dd,
                                           double m)
{
    const double ang_vel = v * r.x;
    const double sqn = ang_vel * ang_vel;
    const double rce = (r.y == r.y) ? rc : r.y;
    const double accel_ang = accel_ang_fast(sqn - sq, r.y, rce);
    const double x = accel_ang * hw;
    const double l = acc + x, rr = acc - x;
    double a_wheel = (fabs(l) < fabs(rr)) ? l : rr;
    if (a_wheel < 0) a_wheel = 0;
    const double dcl = pymin(r.z, a_wheel);
    const double pv = sqrt(v * v + 2 * dcl * dd);
    sq = sqn;
    return pymin(pv, m);
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk-speculative passes: CTA = one path, thread c = chunk c (blockDim.x chunks).
//
// Sweep 1: every thread runs its own chunk from a guessed state (lockstep, all lanes busy).  Fix-up rounds: a chunk whose
// predecessor's end state differs bitwise from the state it started from re-runs from the true state; the re-run stops as
// soon as two consecutive velocities equal the stored ones bitwise (from there the old values ARE the serial values).
// Rounds end when no chunk had to re-run; then, by induction from chunk 0, every value is the serial value.
//
// The kernels are bound by the latency of the dependent fp64 chain of one step, so they are written for residency
// (<= 64 registers: 32 one-warp CTAs per SM) and a short chain: regime tables in shared memory, 32-bit indices, records
// fetched one step ahead into two named buffers (no register rotation), the division's reciprocal from the pre-pass.
// ------------------------------------------------------------------------------------------------------------------
#define CH_INT_MAX 2147483647
#ifndef CH_PF
#define CH_PF 4          // rows the passes prefetch into L1 ahead of the row being loaded into registers
#endif
__device__ __forceinline__ void pf_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// forward chunk of column `col`: edges lo .. lo+len-1; edge e reads slot (e-lo)*NT + col and writes the forward velocity
// of sample e+1 into the slot of edge e+1 (the bottom of the chunk writes row 0 of the next column, or `tail`).
template <bool RERUN>
__device__ __forceinline__ bool fwd_run(const double4* __restrict__ F, const double* __restrict__ RG, double* __restrict__ vf,
                                        int NT, int col, int lo, int len, int last_slot, const int* s_bi, const double* s_acc,
                                        int n_b, double hw, double dd, double& v, double& sq, bool prev_same)
{
    int j = 0;
    while (j + 1 < n_b && s_bi[j + 1] <= lo) j++;
    double acc = s_acc[j];
    int nb_next = (j + 1 < n_b) ? s_bi[j + 1] : CH_INT_MAX;
    int slot = col;                                   // slot of the current edge
    double4 ra = ldg_d4(F + slot), rb;
    double ga = __ldg(RG + slot), gb;
    double olda = 0.0, oldb = 0.0;
    if (RERUN) olda = vf[(len > 1) ? slot + NT : last_slot];
    int e = lo;
    const int hi = lo + len;
    while (true) {
        // ---- buffers a (look-ahead loads never leave the path's row: it is padded by NT slots)
        if (e + CH_PF < hi) { pf_l1(F + slot + CH_PF * NT); pf_l1(RG + slot + CH_PF * NT); }
        rb = ldg_d4(F + slot + NT); gb = __ldg(RG + slot + NT);
        if (RERUN) oldb = vf[(e + 2 < hi) ? slot + 2 * NT : last_slot];
        if (e == nb_next) { j++; acc = s_acc[j]; nb_next = (j + 1 < n_b) ? s_bi[j + 1] : CH_INT_MAX; }
        v = fwd_step(ra, ga, v, sq, acc, hw, dd);
        if (RERUN) {
            const bool same = same_bits(olda, v);
            if (same && prev_same) return true;                         // state equals the old run's: the rest is unchanged
            prev_same = same;
        }
        if (++e >= hi) { vf[last_slot] = v; break; }
        slot += NT;
        vf[slot] = v;
        // ---- buffers b
        if (e + CH_PF < hi) { pf_l1(F + slot + CH_PF * NT); pf_l1(RG + slot + CH_PF * NT); }
        ra = ldg_d4(F + slot + NT); ga = __ldg(RG + slot + NT);
        if (RERUN) olda = vf[(e + 2 < hi) ? slot + 2 * NT : last_slot];
        if (e == nb_next) { j++; acc = s_acc[j]; nb_next = (j + 1 < n_b) ? s_bi[j + 1] : CH_INT_MAX; }
        v = fwd_step(rb, gb, v, sq, acc, hw, dd);
        if (RERUN) {
            const bool same = same_bits(oldb, v);
            if (same && prev_same) return true;
            prev_same = same;
        }
        if (++e >= hi) { vf[last_slot] = v; break; }
        slot += NT;
        vf[slot] = v;
    }
    return false;
}

// Forward pass.  Chunk c = thread c.  vfT: forward velocities in slot order (vfT[slot(e)] = velocity at sample e).
__global__ void __maxnreg__(72) k_fwd_chunked(
    const int* __restrict__ status, const double* __restrict__ cons, double dd, double start_vel, long long RS,
    const int* __restrict__ n_samples, const double4* __restrict__ recF, const double* __restrict__ rg, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, double* __restrict__ vfT, int* __restrict__ rounds_out)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int NT = blockDim.x, c = threadIdx.x;
    const long long b = blockIdx.x;
    double* s_endv = reinterpret_cast<double*>(s_mem);           // [NT] end state of every chunk: v and (v|k|)^2
    double* s_endw = s_endv + NT;
    double* s_usev = s_endw + NT;                                 // [NT] start state every chunk last used
    double* s_usew = s_usev + NT;
    double* s_acc = s_usew + NT;                                  // [E_cap] max_acc of regime j
    int* s_bi = reinterpret_cast<int*>(s_acc + E_cap);            // [E_cap] first sample of regime j
    if (status[b] != ST_OK) return;                               // uniform over the CTA
    const int D = n_samples[b];
    const int steps = D - 1;
    double* vf = vfT + (size_t)b * RS;
    if (c == 0) { vf[0] = start_vel; vf[RS - 1] = start_vel; if (rounds_out) rounds_out[2 * b] = 0; }
    if (steps <= 0) return;
    const int n_b = n_ev[2 * b + 1];
    for (int k = c; k < n_b; k += NT) {
        s_bi[k] = bidx[(size_t)b * E_cap + k];
        s_acc[k] = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + k]];
    }
    const int Lc = chunk_len(steps, NT);
    const int nch = (steps + Lc - 1) / Lc;
    const bool active = c < nch;
    const int lo = c * Lc;
    const int len = (lo + Lc < steps) ? Lc : steps - lo;
    const int last_slot = (lo + len < steps) ? c + 1 : (int)(RS - 1);       // where the velocity of sample lo+len goes
    const double4* F = recF + (size_t)b * RS;
    const double* RG = rg + (size_t)b * RS;
    const double hw = cons[b * 6 + 5] * 0.5;
    __syncthreads();

    // ---- sweep 1: chunk c > 0 starts from the guess "the state-independent caps bind on the two samples before it"
    {
        double v = start_vel, sq = 0.0;
        if (active) {
            if (c > 0) {
                const double vm1 = (lo >= 2) ? F[edge_slot(lo - 2, Lc, NT)].w : start_vel;
                const double4 fm1 = F[edge_slot(lo - 1, Lc, NT)];
                v = fm1.w;
                const double wp = vm1 * fm1.x;
                sq = wp * wp;
            }
            s_usev[c] = v; s_usew[c] = sq;
            fwd_run<false>(F, RG, vf, NT, c, lo, len, last_slot, s_bi, s_acc, n_b, hw, dd, v, sq, false);
        }
        s_endv[c] = v; s_endw[c] = sq;
    }
    __syncthreads();

    // ---- fix-up rounds
    int rounds = 0;
    for (int round = 1; round < NT; round++) {
        bool need = false;
        double in_v = 0.0, in_w = 0.0;
        if (active && c >= round) {
            in_v = s_endv[c - 1]; in_w = s_endw[c - 1];
            need = !(same_bits(in_v, s_usev[c]) && same_bits(in_w, s_usew[c]));
        }
        if (!__syncthreads_or(need)) break;          // also orders this round's reads before its writes
        rounds = round;
        if (need) {
            double v = in_v, sq = in_w;
            const bool merged = fwd_run<true>(F, RG, vf, NT, c, lo, len, last_slot, s_bi, s_acc, n_b, hw, dd, v, sq,
                                              same_bits(in_v, s_usev[c]));
            s_usev[c] = in_v; s_usew[c] = in_w;
            if (!merged) { s_endv[c] = v; s_endw[c] = sq; }
        }
        __syncthreads();
    }
    if (c == 0 && rounds_out) rounds_out[2 * b] = rounds;
}

// backward chunk of column `col`: edges lo+len-1 down to lo; edge e (backward step i = e+1 -> e) reads slot (e-lo)*NT + col
// of the backward records, the reciprocals and the forward velocities, and writes the final velocity of sample e into the
// same slot of voT.
template <bool RERUN>
__device__ __forceinline__ bool bwd_run(const double4* __restrict__ R, const double* __restrict__ RG,
                                        const double* __restrict__ vf, double* __restrict__ voT, int NT, int col, int lo,
                                        int len, const int* s_bi, const double* s_acc, int n_b, double acc0, double hw,
                                        double dd, double& v, double& sq, bool prev_same)
{
    // regime at the chunk start (walking down from D-1): the smallest boundary index > i was the last one applied
    int e = lo + len - 1;                            // current edge; the reference's loop index is i = e + 1
    int j = n_b - 1;
    double acc = acc0;
    while (j >= 0 && s_bi[j] > e + 1) { acc = s_acc[j]; j--; }
    int nb_next = (j >= 0) ? s_bi[j] : -1;
    int slot = (len - 1) * NT + col;
    double4 ra = ldg_d4(R + slot), rb;
    double ga = __ldg(RG + slot), gb;
    double fa = __ldg(vf + slot), fb;
    double olda = 0.0, oldb = 0.0;
    if (RERUN) olda = voT[slot];
    while (true) {
        // ---- buffers a (look-ahead slot clamped to the chunk's first row)
        if (e - CH_PF >= lo) { pf_l1(R + slot - CH_PF * NT); pf_l1(RG + slot - CH_PF * NT); pf_l1(vf + slot - CH_PF * NT); }
        {
            const int sn = (e > lo) ? slot - NT : slot;
            rb = ldg_d4(R + sn); gb = __ldg(RG + sn); fb = __ldg(vf + sn);
            if (RERUN) oldb = voT[sn];
        }
        if (e + 1 == nb_next) { acc = s_acc[j]; j--; nb_next = (j >= 0) ? s_bi[j] : -1; }
        v = bwd_step(ra, ga, v, sq, acc, hw, dd, pymin(fa, ra.w));
        if (RERUN) {
            const bool same = same_bits(olda, v);
            if (same && prev_same) return true;
            prev_same = same;
        }
        voT[slot] = v;
        if (--e < lo) break;
        slot -= NT;
        // ---- buffers b
        if (e - CH_PF >= lo) { pf_l1(R + slot - CH_PF * NT); pf_l1(RG + slot - CH_PF * NT); pf_l1(vf + slot - CH_PF * NT); }
        {
            const int sn = (e > lo) ? slot - NT : slot;
            ra = ldg_d4(R + sn); ga = __ldg(RG + sn); fa = __ldg(vf + sn);
            if (RERUN) olda = voT[sn];
        }
        if (e + 1 == nb_next) { acc = s_acc[j]; j--; nb_next = (j >= 0) ? s_bi[j] : -1; }
        v = bwd_step(rb, gb, v, sq, acc, hw, dd, pymin(fb, rb.w));
        if (RERUN) {
            const bool same = same_bits(oldb, v);
            if (same && prev_same) return true;
            prev_same = same;
        }
        voT[slot] = v;
        if (--e < lo) break;
        slot -= NT;
    }
    return false;
}

// Backward pass.  Thread k walks column nch-1-k (the k-th chunk counted from the end), so states flow from thread k-1 to
// thread k as in the forward kernel.  Reads the forward velocities and writes the final ones, both in slot order
// (velT[RS-1] = end_vel is the last sample); k_untranspose puts them into sample order.  Also accumulates the
// travel-time estimate used to size the time-domain outputs.
__global__ void __maxnreg__(72) k_bwd_chunked(
    const int* __restrict__ status, const double* __restrict__ cons, double dd, double dt, double end_vel,
    long long RS, const int* __restrict__ n_samples, const double4* __restrict__ recR, const double* __restrict__ rg,
    int E_cap, const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const double* __restrict__ vfT, double* __restrict__ velT, float* __restrict__ t_est,
    int* __restrict__ rounds_out)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int NT = blockDim.x, k = threadIdx.x;
    const long long b = blockIdx.x;
    double* s_endv = reinterpret_cast<double*>(s_mem);
    double* s_endw = s_endv + NT;
    double* s_usev = s_endw + NT;
    double* s_usew = s_usev + NT;
    double* s_acc = s_usew + NT;                                  // [E_cap] max_accels[bval[j] + 1] (applied at sample bidx[j])
    int* s_bi = reinterpret_cast<int*>(s_acc + E_cap);
    if (k == 0 && t_est) t_est[b] = 0.f;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int steps = D - 1;
    double* vo = velT + (size_t)b * RS;
    if (k == 0) { vo[RS - 1] = end_vel; if (rounds_out) rounds_out[2 * b + 1] = 0; }
    if (steps <= 0) return;
    const int n_b = n_ev[2 * b + 1];
    for (int q = k; q < n_b; q += NT) {
        s_bi[q] = bidx[(size_t)b * E_cap + q];
        s_acc[q] = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + q] + 1];
    }
    const double acc0 = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + n_b - 1]];
    const int Lc = chunk_len(steps, NT);
    const int nch = (steps + Lc - 1) / Lc;
    const bool active = k < nch;
    const int c = nch - 1 - k;                                     // column (forward chunk index)
    const int lo = c * Lc;
    const int len = (lo + Lc < steps) ? Lc : steps - lo;
    const double4* R = recR + (size_t)b * RS;
    const double* RG = rg + (size_t)b * RS;
    const double* vf = vfT + (size_t)b * RS;
    const double hw = cons[b * 6 + 5] * 0.5;
    __syncthreads();

    // ---- sweep 1: guess v[hi] = min(vel_f[hi], G[hi+1]) (the state-independent part of what step hi+1 produces)
    {
        double v = end_vel, sq = 0.0;
        if (active) {
            if (k > 0) {
                const int hi = lo + len;                                   // sample at the top of this chunk (< D-1)
                const int s1 = edge_slot(hi, Lc, NT);                      // edge hi: backward record of sample hi+1
                const double4 r1 = R[s1];
                v = pymin(vf[s1], r1.w);
                double vp1 = end_vel;
                if (hi + 2 <= D - 1) { const int s2 = edge_slot(hi + 1, Lc, NT); vp1 = pymin(vf[s2], R[s2].w); }
                const double wp = vp1 * r1.x;
                sq = wp * wp;
            }
            s_usev[k] = v; s_usew[k] = sq;
            bwd_run<false>(R, RG, vf, vo, NT, c, lo, len, s_bi, s_acc, n_b, acc0, hw, dd, v, sq, false);
        }
        s_endv[k] = v; s_endw[k] = sq;
    }
    __syncthreads();

    // ---- fix-up rounds (states flow from thread k-1 to thread k, as in the forward kernel)
    int rounds = 0;
    for (int round = 1; round < NT; round++) {
        bool need = false;
        double in_v = 0.0, in_w = 0.0;
        if (active && k >= round) {
            in_v = s_endv[k - 1]; in_w = s_endw[k - 1];
            need = !(same_bits(in_v, s_usev[k]) && same_bits(in_w, s_usew[k]));
        }
        if (!__syncthreads_or(need)) break;
        rounds = round;
        if (need) {
            double v = in_v, sq = in_w;
            const bool merged = bwd_run<true>(R, RG, vf, vo, NT, c, lo, len, s_bi, s_acc, n_b, acc0, hw, dd, v, sq,
                                              same_bits(in_v, s_usev[k]));
            s_usev[k] = in_v; s_usew[k] = in_w;
            if (!merged) { s_endv[k] = v; s_endw[k] = sq; }
        }
        __syncthreads();
    }
    if (k == 0 && rounds_out) rounds_out[2 * b + 1] = rounds;

    // ---- travel-time estimate (single precision is plenty: it only sizes buffers)
    __syncthreads();
    float est = 0.f;
    if (active) {
        const int top = (lo + len < steps) ? c + 1 : (int)(RS - 1);       // slot of the sample above the chunk
        float vprev = (float)vo[top];
        for (int sl = (len - 1) * NT + c; sl >= 0; sl -= NT) {
            const float vcur = (float)vo[sl];
            const float vm = 0.5f * (vprev + vcur);
            est += __fdividef((float)dd, fmaxf(vm, 0.05f) * (float)dt);
            vprev = vcur;
        }
    }
    float* s_f = reinterpret_cast<float*>(s_endv);      // end states are no longer needed
    s_f[k] = est;
    __syncthreads();
    if (k == 0 && t_est) {
        float tot = 0.f;
        for (int q = 0; q < NT; q++) tot += s_f[q];
        t_est[b] = tot;
    }
}

// slot order -> sample order through a 32 x 32 shared-memory tile (both sides coalesced): vel[e] = vT[slot(e)], and the
// last sample from slot RS-1.  grid = (row tiles * column tiles, B), 256 threads.
__global__ void __launch_bounds__(256) k_untranspose(const int* __restrict__ status, const int* __restrict__ n_samples,
                                                     long long D_cap, long long RS, int NT,
                                                     const double* __restrict__ vT, double* __restrict__ vel)
{
    __shared__ double tile[32][33];
    const long long b = blockIdx.y;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int steps = D - 1;
    const double* src = vT + (size_t)b * RS;
    double* dst = vel + (size_t)b * D_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[D - 1] = src[RS - 1];
    if (steps <= 0) return;
    const int Lc = chunk_len(steps, NT);
    const int ctiles = NT >> 5;
    const int rt = blockIdx.x / ctiles, ct = blockIdx.x - rt * ctiles;
    const int s0 = rt * 32, c0 = ct * 32;
    if (s0 >= Lc) return;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
    for (int r = wrp; r < 32; r += 8) {
        const int s = s0 + r;
        tile[r][lane] = (s < Lc) ? src[(size_t)s * NT + c0 + lane] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = wrp; q < 32; q += 8) {
        const int c = c0 + q;
        const int s = s0 + lane;
        const int e = c * Lc + s;
        if (s < Lc && e < steps) dst[e] = tile[lane][q];
    }
}
