/*
 * vap_oracle.c -- CPU restatement of the VexAutonomousPlanner spline -> motion-profile path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (vexautonomousplanner_b200/) may
 * import, link or execute this file; it is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the *checker* / CPU baseline.
 *
 * Parity pin: every function below is checked bit-for-bit (integer outputs and all IEEE
 * +,-,*,/,sqrt,fma arithmetic) against fixtures in tests/golden/ that were produced by running
 * the unmodified Python reference headless (tests/golden/make_golden.py).  Values that go
 * through libm `atan2` / `pow(x,1.5)` (curvature / heading tables) are pinned to <= 4 ulp because
 * numpy's SIMD array loops differ from scalar glibc by an ulp in a few percent of samples.
 *
 * Plain scalar C, one path at a time, IEEE-754 binary64, compiled with -ffp-contract=off so that
 * no multiply-add is fused except where fma() is written explicitly (the places where the
 * reference's BLAS fuses: 1-D np.linalg.norm and the 2x2 matmul).
 *
 * File:line citations are relative to /root/reference/src/.
 *
 * Layouts (shared with include/vap.h):
 *   node_attr[N][12] : x, y, turn_deg, wait_time, max_velocity, max_acceleration,
 *                      tangent_x, tangent_y, incoming_magnitude, outgoing_magnitude, rot_cos, rot_sin
 *   node_flags[N]    : bit0 is_reverse_node, bit1 stop, bit2 tangent is not None
 *   ap_attr[A][4]    : t, wait_time, max_velocity, max_acceleration ; ap_flags[A] bit1 stop
 *   cons[6]          : max_vel, max_acc, max_dec, friction_coef, max_jerk, track_width
 *   seg[G][6][2]     : per global segment g (between node g and g+1): p0,p1,d0,d1,dd0,dd1
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NA 12
enum { A_X = 0, A_Y, A_TURN, A_WAIT, A_MAXVEL, A_MAXACC, A_TX, A_TY, A_INMAG, A_OUTMAG, A_RCOS, A_RSIN };
#define F_REVERSE 1
#define F_STOP 2
#define F_TANGENT 4
#define APA 4
enum { P_T = 0, P_WAIT, P_MAXVEL, P_MAXACC };

#define ORA_ERR_FALSE (-1)      /* reference returns False */
#define ORA_ERR_INDEX (-2)      /* reference raises IndexError */
#define ORA_ERR_VALUE (-3)      /* reference raises ValueError */
#define ORA_ERR_CAPACITY (-4)   /* caller buffer too small (not a reference condition) */

static const double PI = 3.141592653589793;

/* 0: x**2 evaluated as libm pow(x, 2.0) -- what CPython / numpy scalars do (reference-faithful);
 * 1: x**2 evaluated as the correctly rounded product x*x (what the CUDA engine does).
 * glibc's pow is not correctly rounded: pow(x,2.0) != x*x for ~0.08 % of doubles.             */
static int g_sq_mode = 0;
void ora_set_sq_mode(int m) { g_sq_mode = m; }
int ora_get_sq_mode(void) { return g_sq_mode; }
/* the exponent is read through a volatile so that gcc cannot fold pow(x, 2.0) into x*x */
static volatile double g_two = 2.0;
static inline double sq(double x) { return g_sq_mode ? x * x : pow(x, g_two); }

/* np.linalg.norm of a length-2 vector: OpenBLAS dot kernel fuses the second product (A.9) */
static inline double norm1d(double x, double y) { return sqrt(fma(y, y, x * x)); }
/* np.linalg.norm(M, axis=1): plain sqrt(x*x + y*y) */
static inline double norm_ax1(double x, double y) { return sqrt(x * x + y * y); }
/* Python builtin min/max on floats: keeps the first unless a later one compares strictly smaller/larger */
static inline double pymin(double a, double b) { return (b < a) ? b : a; }
static inline double pymax(double a, double b) { return (b > a) ? b : a; }
/* Python float % for positive divisor (also numpy remainder) */
static inline double pymod(double x, double y)
{
    double m = fmod(x, y);
    if (m != 0.0) {
        if ((y < 0) != (m < 0)) m += y;
    } else {
        m = copysign(0.0, y);
    }
    return m;
}
static inline double frac1(double t) { return pymod(t, 1.0); }

/* ------------------------------------------------------------------------------------------
 * S0a. QuinticHermiteSpline.fit for one run of control points (quintic_hermite_spline.py:30-219,
 *      :719-736) with the manager's tangent list (spline_manager.py:65-77).
 *      x,y: n control points.  set_in/set_out: per-node user tangents (has[i] != 0).
 *      start_t / end_t: boundary tangents or NULL.  Outputs seg[(n-1)*12], seglen[n-1], params[n].
 * ------------------------------------------------------------------------------------------ */
static int fit_run(int n, const double* x, const double* y, const int* has, const double* set_in,
                   const double* set_out, const double* start_t, const double* end_t, double* seg,
                   double* seglen, double* params)
{
    if (n < 2) return ORA_ERR_FALSE;
    double* dist = (double*)malloc(sizeof(double) * (size_t)(n - 1));
    double* fd = (double*)calloc((size_t)n * 2, sizeof(double));
    double* sd = (double*)calloc((size_t)n * 2, sizeof(double));
    /* _compute_parameters :719-736 */
    double c = 0.0;
    params[0] = 0.0;
    for (int i = 0; i < n - 1; i++) {
        dist[i] = norm_ax1(x[i + 1] - x[i], y[i + 1] - y[i]);
        c = c + dist[i];           /* np.cumsum: sequential */
        params[i + 1] = c;
    }
    double clast = params[n - 1];
    if (clast == 0.0) {
        /* np.linspace(0, n-1, n) */
        double step = (double)(n - 1) / (double)(n - 1);
        for (int i = 0; i < n; i++) params[i] = (double)i * step + 0.0;
        params[n - 1] = (double)(n - 1);
    } else {
        for (int i = 0; i < n; i++) params[i] = params[i] * (double)(n - 1) / clast;
    }
    /* _compute_derivatives :149-219 (scale_factor == 1) */
    for (int i = 0; i < n; i++) {
        if (i == 0) {
            double cx = x[1] - x[0], cy = y[1] - y[0];
            if (n == 2 && end_t) { fd[0] = cx * 1; fd[1] = cy * 1; }
            else { fd[0] = cx * 1 / dist[0]; fd[1] = cy * 1 / dist[0]; }
        } else if (i == n - 1) {
            double cx = x[n - 1] - x[n - 2], cy = y[n - 1] - y[n - 2];
            if (n == 2 && start_t) { fd[2 * i] = cx * 1; fd[2 * i + 1] = cy * 1; }
            else { fd[2 * i] = cx * 1 / dist[n - 2]; fd[2 * i + 1] = cy * 1 / dist[n - 2]; }
        } else {
            double px = (x[i] - x[i - 1]) / dist[i - 1], py = (y[i] - y[i - 1]) / dist[i - 1];
            double nx = (x[i + 1] - x[i]) / dist[i], ny = (y[i + 1] - y[i]) / dist[i];
            fd[2 * i] = (px + nx) * 1 / 2;
            fd[2 * i + 1] = (py + ny) * 1 / 2;
        }
    }
    for (int i = 1; i < n - 1; i++) {
        double avg = (dist[i - 1] + dist[i]) / 2;
        double den = avg * 0.5;
        sd[2 * i] = (fd[2 * (i + 1)] - fd[2 * (i - 1)]) / den;
        sd[2 * i + 1] = (fd[2 * (i + 1) + 1] - fd[2 * (i - 1) + 1]) / den;
    }
    /* segments :76-127 */
    for (int i = 0; i < n - 1; i++) {
        double* s = seg + (size_t)i * 12;
        double L = norm1d(x[i + 1] - x[i], y[i + 1] - y[i]);
        seglen[i] = L;
        s[0] = x[i]; s[1] = y[i]; s[2] = x[i + 1]; s[3] = y[i + 1];
        if (L > 0) {
            double L2 = sq(L);
            s[4] = fd[2 * i] * L; s[5] = fd[2 * i + 1] * L;
            s[6] = fd[2 * (i + 1)] * L; s[7] = fd[2 * (i + 1) + 1] * L;
            s[8] = sd[2 * i] * L2; s[9] = sd[2 * i + 1] * L2;
            s[10] = sd[2 * (i + 1)] * L2; s[11] = sd[2 * (i + 1) + 1] * L2;
            if (has[i]) { s[4] = set_out[2 * i]; s[5] = set_out[2 * i + 1]; }
            if (has[i + 1]) { s[6] = set_in[2 * (i + 1)]; s[7] = set_in[2 * (i + 1) + 1]; }
        } else {
            s[4] = fd[2 * i]; s[5] = fd[2 * i + 1];
            s[6] = fd[2 * (i + 1)]; s[7] = fd[2 * (i + 1) + 1];
            s[8] = sd[2 * i]; s[9] = sd[2 * i + 1];
            s[10] = sd[2 * (i + 1)]; s[11] = sd[2 * (i + 1) + 1];
        }
    }
    /* :129-132 + :543-590 -- the start tangent lands on the LAST segment's row 2 (load-bearing bug) */
    double* last = seg + (size_t)(n - 2) * 12;
    if (start_t) { last[4] = start_t[0]; last[5] = start_t[1]; }
    if (end_t) { last[6] = end_t[0]; last[7] = end_t[1]; }
    free(dist); free(fd); free(sd);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * S0b. QuinticHermiteSplineManager.build_path (spline_manager.py:42-172).
 *   Returns the number of splines S (>=1) or a negative error.
 *   first_node[S+1]: node index where each spline starts (last entry = N-1).
 *   param_end[S]: spline.parameters[-1].  params_concat: every spline's parameter array, concatenated
 *   (length N - 1 + S).
 * ------------------------------------------------------------------------------------------ */
int ora_build_path(int n, const double* na, const int* nf, double* seg, int* first_node,
                   double* param_end, double* seglen, double* params_concat)
{
    if (n < 2) return ORA_ERR_FALSE;
    double* xs = (double*)malloc(sizeof(double) * (size_t)n);
    double* ys = (double*)malloc(sizeof(double) * (size_t)n);
    double* tin = (double*)malloc(sizeof(double) * (size_t)n * 2);
    double* tout = (double*)malloc(sizeof(double) * (size_t)n * 2);
    int* has = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const double* a = na + (size_t)i * NA;
        xs[i] = a[A_X]; ys[i] = a[A_Y];
        has[i] = (nf[i] & F_TANGENT) != 0;
        /* :67-73  node.tangent * magnitude */
        tin[2 * i] = a[A_TX] * a[A_INMAG]; tin[2 * i + 1] = a[A_TY] * a[A_INMAG];
        tout[2 * i] = a[A_TX] * a[A_OUTMAG]; tout[2 * i + 1] = a[A_TY] * a[A_OUTMAG];
    }
    int S = 0, cur = 0, pc = 0, rc = 0;
    int have_start = 0;
    double start_t[2] = {0, 0};
    first_node[0] = 0;
    for (int i = 1; i < n && rc == 0; i++) {
        const double* a = na + (size_t)i * NA;
        int rev = (nf[i] & F_REVERSE) != 0;
        int turn = a[A_TURN] != 0;
        if (!(rev || turn || i == n - 1)) continue;
        double end_t[2];
        int have_end = 0;
        double use_start[2] = {start_t[0], start_t[1]};
        int use_have_start = have_start;
        have_start = 0;   /* :79-81 */
        if (rev || turn) {
            double dxp = xs[i] - xs[i - 1], dyp = ys[i] - ys[i - 1];
            double prev_len = norm1d(dxp, dyp);
            if (i >= n - 1) { rc = ORA_ERR_INDEX; break; }   /* points[i + 1] :97 */
            double dxn = xs[i + 1] - xs[i], dyn = ys[i + 1] - ys[i];
            double next_len = norm1d(dxn, dyn);
            double sp = prev_len > 0 ? 1.0 / prev_len : 1.0;
            double sn = next_len > 0 ? 1.0 / next_len : 1.0;
            double pvx = dxp * sp, pvy = dyp * sp;
            double nvx = dxn * sn, nvy = dyn * sn;
            double min_len = pymin(prev_len, next_len);
            if (turn) {
                double c = a[A_RCOS], s = a[A_RSIN];
                /* rotation_matrix @ prev_vector: BLAS fuses the first product into the second */
                double ntx = fma(c, pvx, (-s) * pvy);
                double nty = fma(s, pvx, c * pvy);
                ntx = ntx * min_len; nty = nty * min_len;
                pvx = pvx * min_len; pvy = pvy * min_len;
                if (has[i]) {
                    pvx = a[A_TX] * a[A_INMAG]; pvy = a[A_TY] * a[A_INMAG];
                    /* tangent @ rotation_matrix * -1 */
                    double vx = a[A_TX], vy = a[A_TY];
                    double r0 = fma(vy, s, vx * c);
                    double r1 = fma(vy, c, vx * (-s));
                    ntx = (r0 * -1) * a[A_OUTMAG]; nty = (r1 * -1) * a[A_OUTMAG];
                }
                end_t[0] = pvx; end_t[1] = pvy; have_end = 1;
                start_t[0] = ntx; start_t[1] = nty; have_start = 1;
            } else {
                double dx = pvx - nvx, dy = pvy - nvy;
                double dn = norm1d(dx, dy);
                if (dn > 0) { dx = dx / dn; dy = dy / dn; }
                dx = dx * min_len; dy = dy * min_len;
                if (has[i]) { dx = a[A_TX] * a[A_INMAG]; dy = a[A_TY] * a[A_INMAG]; }
                end_t[0] = dx; end_t[1] = dy; have_end = 1;
                start_t[0] = -1 * dx; start_t[1] = -1 * dy; have_start = 1;
                if (has[i]) {
                    start_t[0] = (-1 * a[A_TX]) * a[A_OUTMAG];
                    start_t[1] = (-1 * a[A_TY]) * a[A_OUTMAG];
                }
            }
        }
        int m = i - cur + 1;
        int r = fit_run(m, xs + cur, ys + cur, has + cur, tin + 2 * cur, tout + 2 * cur,
                        use_have_start ? use_start : NULL, have_end ? end_t : NULL,
                        seg + (size_t)cur * 12, seglen + cur, params_concat + pc);
        if (r < 0) { rc = r; break; }
        param_end[S] = params_concat[pc + m - 1];
        pc += m;
        S++;
        first_node[S] = i;
        if ((rev || turn) && i < n - 1) cur = i;
    }
    free(xs); free(ys); free(tin); free(tout); free(has);
    return rc < 0 ? rc : S;
}

/* ------------------------------------------------------------------------------------------
 * Evaluation.  which: 0 point, 1 first derivative, 2 second derivative.
 * Spline level: quintic_hermite_spline.py:506-541 (normalise), :288-416 (basis), :221-251/:473-504.
 * ------------------------------------------------------------------------------------------ */
static void basis(int which, double t, double* H)
{
    double t2 = t * t, t3 = t2 * t;
    if (which == 0) {
        double t4 = t3 * t, t5 = t4 * t;
        H[0] = 1 - 10 * t3 + 15 * t4 - 6 * t5;
        H[1] = 10 * t3 - 15 * t4 + 6 * t5;
        H[2] = t - 6 * t3 + 8 * t4 - 3 * t5;
        H[3] = -4 * t3 + 7 * t4 - 3 * t5;
        H[4] = 0.5 * t2 - 1.5 * t3 + 1.5 * t4 - 0.5 * t5;
        H[5] = 0.5 * t3 - t4 + 0.5 * t5;
    } else if (which == 1) {
        double t4 = t3 * t;
        H[0] = -30 * t2 + 60 * t3 - 30 * t4;
        H[1] = 30 * t2 - 60 * t3 + 30 * t4;
        H[2] = 1 - 18 * t2 + 32 * t3 - 15 * t4;
        H[3] = -12 * t2 + 28 * t3 - 15 * t4;
        H[4] = t - 4.5 * t2 + 6 * t3 - 2.5 * t4;
        H[5] = 1.5 * t2 - 4 * t3 + 2.5 * t4;
    } else {
        H[0] = -60 * t + 180 * t2 - 120 * t3;
        H[1] = 60 * t - 180 * t2 + 120 * t3;
        H[2] = -36 * t + 96 * t2 - 60 * t3;
        H[3] = -24 * t + 84 * t2 - 60 * t3;
        H[4] = 1 - 9 * t + 18 * t2 - 10 * t3;
        H[5] = 3 * t - 12 * t2 + 10 * t3;
    }
}

void ora_eval_spline(int nseg, double pend, const double* seg, int which, double t, double* out)
{
    double t_min = 0.0, t_max = pend;
    double tt = pymax(t_min, pymin(t, t_max));
    int idx = (int)((tt - t_min) / 1.0);
    if (idx == nseg) idx = nseg - 1;
    double seg_start = t_min + idx * 1.0;
    double u = (tt - seg_start) / 1.0;
    double H[6];
    basis(which, u, H);
    const double* s = seg + (size_t)idx * 12;
    double ax = 0.0, ay = 0.0;
    for (int i = 0; i < 6; i++) { ax += H[i] * s[2 * i]; ay += H[i] * s[2 * i + 1]; }
    out[0] = ax; out[1] = ay;
}

/* manager level: spline_manager.py:243-275 then spline-level eval */
void ora_eval(int S, const int* first_node, const double* param_end, const double* seg, int which,
              double t, double* out)
{
    int k = 0;
    for (k = 0; k < S; k++) {
        double seg_end = (double)first_node[k + 1];   /* cumulative + num_points - 1 */
        if (t <= seg_end || k == S - 1) break;
    }
    double local = t - (double)first_node[k];
    ora_eval_spline(first_node[k + 1] - first_node[k], param_end[k], seg + (size_t)first_node[k] * 12,
                    which, local, out);
}

/* Spline.get_heading / get_curvature (spline.py:48-80) on the manager (_get_heading/_get_curvature
 * spline_manager.py:348-418 have the same arithmetic except the zero test). */
double ora_exact_heading(int S, const int* fn, const double* pe, const double* seg, double t)
{
    double d[2];
    ora_eval(S, fn, pe, seg, 1, t, d);
    return atan2(d[1], d[0]);
}
double ora_exact_curvature(int S, const int* fn, const double* pe, const double* seg, double t)
{
    double d1[2], d2[2];
    ora_eval(S, fn, pe, seg, 1, t, d1);
    ora_eval(S, fn, pe, seg, 2, t, d2);
    double s2 = sq(d1[0]) + sq(d1[1]);
    if (s2 < 1e-10) return 0.0;
    return (d1[0] * d2[1] - d1[1] * d2[0]) / pow(s2, 1.5);
}

/* ------------------------------------------------------------------------------------------
 * S1. build_lookup_table (spline_manager.py:426-475).  lut_d/lut_t have samples*S entries.
 * ------------------------------------------------------------------------------------------ */
double ora_build_lut(int S, const int* first_node, const double* param_end, const double* seg,
                     int samples, double* lut_d, double* lut_t)
{
    double current = 0.0, prev_param = 0.0;
    double* mag = (double*)malloc(sizeof(double) * (size_t)samples);
    double* lp = (double*)malloc(sizeof(double) * (size_t)samples);
    for (int k = 0; k < S; k++) {
        double p0 = 0.0, p1 = param_end[k];
        int nseg = first_node[k + 1] - first_node[k];
        const double* sg = seg + (size_t)first_node[k] * 12;
        /* np.linspace(p0, p1, samples) */
        double step = (p1 - p0) / (double)(samples - 1);
        for (int j = 0; j < samples; j++) lp[j] = (double)j * step + p0;
        lp[samples - 1] = p1;
        double dt = lp[1] - lp[0];
        for (int j = 0; j < samples; j++) {
            double d[2];
            ora_eval_spline(nseg, p1, sg, 1, lp[j], d);
            mag[j] = norm_ax1(d[0], d[1]);
        }
        double* od = lut_d + (size_t)k * samples;
        double* ot = lut_t + (size_t)k * samples;
        double c = 0.0;
        od[0] = 0.0 + current;
        for (int j = 1; j < samples; j++) {
            double inc = (mag[j - 1] + mag[j]) * 0.5 * dt;
            c = (j == 1) ? inc : c + inc;
            od[j] = c + current;
        }
        for (int j = 0; j < samples; j++) ot[j] = lp[j] + prev_param;
        current = od[samples - 1];
        prev_param += p1 - p0;
    }
    free(mag); free(lp);
    return current;
}

/* ------------------------------------------------------------------------------------------
 * S2. precompute_path_properties (spline_manager.py:477-548).  P = spn * n entries.
 * ------------------------------------------------------------------------------------------ */
static inline double prop_param(long j, long P, int n, double step)
{
    return (j == P - 1) ? (double)(n - 1) : (double)j * step + 0.0;
}
void ora_build_props(int n, int S, const int* first_node, const double* param_end, const double* seg,
                     int spn, double* kap, double* th)
{
    long P = (long)spn * n;
    double step = (double)(n - 1) / (double)(P - 1);
    for (long j = 0; j < P; j++) {
        double t = prop_param(j, P, n, step);
        double d1[2], d2[2];
        ora_eval(S, first_node, param_end, seg, 1, t, d1);
        ora_eval(S, first_node, param_end, seg, 2, t, d2);
        double s2 = d1[0] * d1[0] + d1[1] * d1[1];      /* array ** 2 is an exact square */
        double num = d1[0] * d2[1] - d1[1] * d2[0];
        kap[j] = (s2 >= 1e-10) ? num / pow(s2, 1.5) : 0.0;
        th[j] = atan2(d1[1], d1[0]);
    }
}

/* distance_to_time (spline_manager.py:291-318) */
double ora_distance_to_time(long Q, const double* lut_d, const double* lut_t, double total, int n, double d)
{
    if (d <= 0) return 0.0;
    if (d >= total) return (double)(n - 1);
    long lo = 0, hi = Q;   /* np.searchsorted side='left' */
    while (lo < hi) {
        long mid = lo + (hi - lo) / 2;
        if (lut_d[mid] < d) lo = mid + 1; else hi = mid;
    }
    long idx = lo;
    if (idx == 0) return lut_t[0];
    double d0 = lut_d[idx - 1], d1 = lut_d[idx], t0 = lut_t[idx - 1], t1 = lut_t[idx];
    return t0 + (t1 - t0) * (d - d0) / (d1 - d0);
}

/* _interpolate_property (spline_manager.py:550-580); parameters are the analytic linspace */
double ora_snap(long P, int n, const double* vals, double t)
{
    double step = (double)(n - 1) / (double)(P - 1);
    long lo = 0, hi = P;
    while (lo < hi) {
        long mid = lo + (hi - lo) / 2;
        if (prop_param(mid, P, n, step) < t) lo = mid + 1; else hi = mid;
    }
    long idx = lo;
    if (idx == 0) return vals[0];
    if (idx >= P) return vals[P - 1];
    double t0 = prop_param(idx - 1, P, n, step), t1 = prop_param(idx, P, n, step);
    if (frac1(t0) != frac1(t1)) return (frac1(t) > 0.5) ? vals[idx - 1] : vals[idx];
    double v0 = vals[idx - 1], v1 = vals[idx];
    return v0 + (v1 - v0) * (t - t0) / (t1 - t0);
}

/* ------------------------------------------------------------------------------------------
 * S3. forward_backward_pass -- sampling + event stage (motion_profile_generator.py:80-176).
 * Returns D (number of distance samples incl. the final one) or ORA_ERR_CAPACITY.
 * ------------------------------------------------------------------------------------------ */
long ora_dist_sample(int n, const double* na, const int* nf, int A, const double* apa, const int* apf,
                     const double* cons, double dd, double end_vel, long Q, const double* lut_d,
                     const double* lut_t, double total, long P, const double* kap, const double* th,
                     long cap, double* t_out, double* k_out, double* h_out, double* v_out,
                     double* max_accels, int* n_acc_out, long* bidx, int* bval, int* n_b_out)
{
    double V = cons[0], Acc = cons[1];
    double current = 0.0, prev_t = 0.0;
    int node_num = 0, action_idx = 0;
    double max_velocity = V;
    int n_acc = 0, n_b = 0;
    max_accels[n_acc++] = na[A_MAXACC] > 0 ? na[A_MAXACC] : Acc;
    if (na[A_MAXVEL] > 0) max_velocity = na[A_MAXVEL];
    bidx[n_b] = 0; bval[n_b] = 0; n_b++;
    long i = 0;
    double t_end = ora_distance_to_time(Q, lut_d, lut_t, total, n, total);
    while (current < total) {
        if (i + 1 >= cap) return ORA_ERR_CAPACITY;
        double t = ora_distance_to_time(Q, lut_d, lut_t, total, n, current);
        t_out[i] = t;
        k_out[i] = ora_snap(P, n, kap, t);
        h_out[i] = ora_snap(P, n, th, t);
        v_out[i] = max_velocity;
        current += dd;
        if (frac1(prev_t) > frac1(t) && t < t_end) {
            node_num += 1;
            const double* a = na + (size_t)node_num * NA;
            if (nf[node_num] & F_STOP) v_out[i] = 0.01;
            max_velocity = a[A_MAXVEL] > 0 ? a[A_MAXVEL] : V;
            max_accels[n_acc++] = a[A_MAXACC] > 0 ? a[A_MAXACC] : Acc;
            if (node_num < n - 1) {
                if (bidx[n_b - 1] == i) bval[n_b - 1] = n_acc - 1;
                else { bidx[n_b] = i; bval[n_b] = n_acc - 1; n_b++; }
            }
        }
        if (action_idx < A && prev_t < apa[action_idx * APA + P_T] && t >= apa[action_idx * APA + P_T]) {
            const double* p = apa + (size_t)action_idx * APA;
            max_velocity = p[P_MAXVEL] > 0 ? p[P_MAXVEL] : V;
            if (apf[action_idx] & F_STOP) v_out[i] = 0.01;
            max_accels[n_acc++] = p[P_MAXACC] > 0 ? p[P_MAXACC] : Acc;
            if (bidx[n_b - 1] == i) bval[n_b - 1] = n_acc - 1;
            else { bidx[n_b] = i; bval[n_b] = n_acc - 1; n_b++; }
            action_idx += 1;
        }
        i += 1;
        prev_t = t;
    }
    v_out[i] = end_vel;
    t_out[i] = t_end;
    h_out[i] = ora_snap(P, n, th, t_end);
    k_out[i] = ora_snap(P, n, kap, t_end);
    max_accels[n_acc++] = Acc;
    *n_acc_out = n_acc;
    *n_b_out = n_b;
    return i + 1;
}

/* ------------------------------------------------------------------------------------------
 * S4 + S5. forward / backward passes (motion_profile_generator.py:188-314).  vel is in/out.
 * ------------------------------------------------------------------------------------------ */
static inline double wheel_accel(double acc, double ang, double w)
{   /* Constraints.max_accels_at_turn :52-59 */
    double l = acc + ang * w / 2;
    double r = acc - ang * w / 2;
    return (fabs(l) < fabs(r)) ? l : r;
}
static inline double max_speed_at_curvature(double V, double w, double curvature)
{   /* :23-33 */
    if (fabs(curvature) < 1e-6) return V;
    double m = ((2 * V / w) * V) / (fabs(curvature) * V + (2 * V / w));
    return pymin(m, V);
}

void ora_fwd_bwd(long D, const double* kap, const double* th, double* vel, const double* cons, double dd,
                 const double* max_accels, int n_b, const long* bidx, const int* bval, double start_vel,
                 double end_vel, int stop_after_forward)
{
    double V = cons[0], acc = cons[1], dec = cons[2], w = cons[5];
    double max_angular_vel = 2 * V / w;
    double max_angular_accel = 2 * acc / w;
    vel[0] = start_vel;
    double prev_ang_vel = 0.0;
    int b = 0;
    for (long i = 0; i < D - 1; i++) {
        if (b < n_b && bidx[b] == i) { acc = max_accels[bval[b]]; dec = max_accels[bval[b]]; b++; }
        double v = vel[i], k = kap[i], ak = fabs(k);
        double ang_vel = v * ak;
        double vlim, a;
        if (ak < 1e-6) { vlim = V; a = acc; }
        else {
            double dth = th[i + 1] - th[i];
            double accel_ang = (sq(ang_vel) - sq(prev_ang_vel)) / (2 * fabs(dth));
            double v_ang = max_angular_vel / ak;
            double v_kin = 2 * V / (w * ak + 2);
            double v_curve = max_speed_at_curvature(V, w, ak);
            vlim = pymin(pymin(v_ang, v_kin), v_curve);
            double a_ang = max_angular_accel / ak;
            double a_kin = 2 * acc / (w * ak + 2);
            double a_wheel = wheel_accel(acc, fabs(accel_ang), w);
            if (a_wheel < 0) a_wheel = 0;
            a = pymin(pymin(pymin(a_ang, a_kin), a_wheel), acc);
        }
        double next_vel = pymin(vlim, sqrt(sq(v) + 2 * a * dd));
        vel[i + 1] = pymin(vel[i + 1], next_vel);
        prev_ang_vel = ang_vel;
        vel[i + 1] = pymin(vel[i + 1], fabs(V / (1 + (w * ak / 2))));
    }
    if (stop_after_forward) return;
    vel[D - 1] = end_vel;
    prev_ang_vel = 0.0;
    b = n_b - 1;
    for (long i = D - 1; i > 0; i--) {
        while (b >= 0 && bidx[b] > i) b--;
        if (b >= 0 && bidx[b] == i) { acc = max_accels[bval[b] + 1]; b--; }
        double v = vel[i], k = kap[i], ak = fabs(k);
        double ang_vel = v * ak;
        double vlim, dcl;
        if (ak < 1e-6) { vlim = V; dcl = dec; }
        else {
            double dth = th[i - 1] - th[i];
            double accel_ang = (sq(ang_vel) - sq(prev_ang_vel)) / (2 * fabs(dth));
            double v_ang = max_angular_vel / ak;
            double v_kin = 2 * V / (w * ak + 2);
            double v_curve = max_speed_at_curvature(V, w, k);
            vlim = pymin(pymin(v_ang, v_kin), v_curve);
            double d_ang = max_angular_accel / ak;
            double d_kin = 2 * dec / (w * ak + 2);
            double a_wheel = wheel_accel(acc, accel_ang, w);
            if (a_wheel < 0) a_wheel = 0;
            dcl = pymin(pymin(pymin(d_ang, d_kin), a_wheel), dec);
        }
        double pv = sqrt(sq(v) + 2 * dcl * dd);
        pv = pymin(pymin(pv, vel[i - 1]), vlim);
        vel[i - 1] = pv;
        prev_ang_vel = ang_vel;
        vel[i - 1] = pymin(vel[i - 1], fabs(V / (1 + (w * ak / 2))));
    }
}

/* ------------------------------------------------------------------------------------------
 * one_dim_mp_generator.generate_trapezoidal_profile (:4-69) and motion_profile_angle
 * (motion_profile_generator.py:319-346).  Return the sample count K, or ORA_ERR_CAPACITY.
 * ------------------------------------------------------------------------------------------ */
long ora_trapezoid(double max_velocity, double max_acceleration, double total_distance, double time_step,
                   long cap, double* vout)
{
    double ttm = max_velocity / max_acceleration;
    double dist_accel = 0.5 * max_acceleration * sq(ttm);
    double total_time;
    if (2 * dist_accel > total_distance) {
        ttm = sqrt(total_distance / max_acceleration);
        max_velocity = max_acceleration * ttm;
        total_time = 2 * ttm;
    } else {
        double dc = total_distance - 2 * dist_accel;
        double tc = dc / max_velocity;
        total_time = 2 * ttm + tc;
    }
    /* np.arange(0, total_time + time_step, time_step): len = ceil((stop-start)/step), values i*step */
    double stop = total_time + time_step;
    long K = (long)ceil((stop - 0.0) / time_step);
    if (K < 0) K = 0;
    if (K > cap) return ORA_ERR_CAPACITY;
    for (long i = 0; i < K; i++) {
        double t = 0.0 + (double)i * time_step;
        double v;
        if (t <= ttm) v = max_acceleration * t;
        else if (t <= total_time - ttm) v = max_velocity;
        else {
            double tid = t - (total_time - ttm);
            v = max_velocity - max_acceleration * tid;
        }
        vout[i] = v;
    }
    return K;
}

long ora_motion_profile_angle(double angle, double V, double A, double w, double dt, long cap,
                              double* heads, double* omegas)
{
    double arc = fabs(angle) * w / 2;
    long K = ora_trapezoid(V, A, arc, dt, cap, omegas /* scratch: velocities */);
    if (K < 0) return K;
    double accum = 0.0;
    double sgn = angle > 0 ? -1.0 : 1.0;
    for (long i = 0; i < K; i++) {
        double v = omegas[i];
        double cur = accum / (w / 2);
        heads[i] = cur * sgn;
        accum += v * dt;
    }
    for (long i = K - 1; i >= 1; i--) omegas[i] = (heads[i] - heads[i - 1]) / dt;
    if (K > 0) omegas[0] = 0.0;
    return K;
}

/* lerp (motion_profile_generator.py:349-386) on xs[i] = i*dd (:484), cache=None */
double ora_lerp_uniform(double x, double dd, long D, const double* ys)
{
    /* np.searchsorted(xs, x, side='right') - 1 with xs[i] = fl(i*dd) */
    long lo = 0, hi = D;
    while (lo < hi) {
        long mid = lo + (hi - lo) / 2;
        if ((double)mid * dd <= x) lo = mid + 1; else hi = mid;
    }
    long idx = lo - 1;
    if (idx < 0) return ys[0];
    if (idx >= D - 1) return ys[D - 1];
    double x0 = (double)idx * dd, x1 = (double)(idx + 1) * dd;
    double y0 = ys[idx], y1 = ys[idx + 1];
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

/* ------------------------------------------------------------------------------------------
 * S6. generate_motion_profile after the velocity passes (motion_profile_generator.py:414-628).
 * Outputs (capacity cap): times, positions, linear_vels, accelerations, headings, angular_vels, x, y.
 * nodes_map gets the caller's trailing len(times) appended (gui/path.py:342).
 * Returns T or a negative error.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    long cap, T;
    double *times, *pos, *lin, *acc, *head, *ang, *x, *y;
} out_t;

static int push(out_t* o, double t, double p, double l, double a, double h, double w, double x, double y)
{
    if (o->T >= o->cap) return 0;
    long i = o->T++;
    o->times[i] = t; o->pos[i] = p; o->lin[i] = l; o->acc[i] = a; o->head[i] = h; o->ang[i] = w;
    o->x[i] = x; o->y[i] = y;
    return 1;
}

static int do_wait(out_t* o, double wait, double* current_time, double dt)
{   /* handle_wait :509-518 */
    long steps = (long)(wait / dt);
    double h = o->head[o->T - 1], x = o->x[o->T - 1], y = o->y[o->T - 1];
    for (long i = 0; i < steps; i++)
        if (!push(o, *current_time + (double)i * dt, 0.0, 0.0, 0.0, h, 0.0, x, y)) return 0;
    *current_time = *current_time + (double)steps * dt;
    return 1;
}

long ora_profile(int n, const double* na, const int* nf, int A, const double* apa, const int* apf,
                 const double* cons, double dt, double dd, int S, const int* first_node,
                 const double* param_end, const double* seg, long Q, const double* lut_d,
                 const double* lut_t, double total, long P, const double* kap, const double* th, long D,
                 const double* vel, long cap, double* times, double* pos, double* lin, double* acc,
                 double* head, double* ang, double* xo, double* yo, long* nodes_map, int* n_nm,
                 long* actions_map, int* n_am)
{
    out_t o = {cap, 0, times, pos, lin, acc, head, ang, xo, yo};
    double V = cons[0], max_acc = cons[1], max_dec = cons[2], w = cons[5];
    double current_time = 0.0, current_pos = 0.0, current_vel = vel[0];
    int is_reversed = 0;
    int nm = 0, am = 0;
    nodes_map[nm++] = 0;
    if (nf[0] & F_REVERSE) is_reversed = !is_reversed;
    if (na[A_TURN] != 0) return ORA_ERR_INDEX;     /* headings[-1] on an empty list :440 */
    if (na[A_WAIT] > 0) {   /* :459-476 */
        long steps = (long)(na[A_WAIT] / dt);
        double h = -1 * ora_snap(P, n, th, 0.0);
        if (is_reversed) h -= PI;
        if (h > PI) h -= 2 * PI;
        if (h < -PI) h += 2 * PI;
        double p0[2];
        ora_eval(S, first_node, param_end, seg, 0, 0.0, p0);
        for (long i = 0; i < steps; i++)
            if (!push(&o, current_time + (double)i * dt, 0.0, 0.0, 0.0, h, 0.0, p0[0], p0[1])) return ORA_ERR_CAPACITY;
        current_time += (double)steps * dt;
    }
    double prev_t = 0.0;
    int action_idx = 0, node_idx = 0;
    double end_param = ora_distance_to_time(Q, lut_d, lut_t, total, n, total);
    long tcap = 1 << 16;
    double* ins_h = (double*)malloc(sizeof(double) * (size_t)tcap);
    double* ins_w = (double*)malloc(sizeof(double) * (size_t)tcap);
    long rc = 0;
    while (current_pos < total) {
        double t = ora_distance_to_time(Q, lut_d, lut_t, total, n, current_pos);
        if (frac1(t) < frac1(prev_t) && t < end_param) {
            nodes_map[nm++] = o.T;
            node_idx += 1;
            if (node_idx >= n) { rc = ORA_ERR_INDEX; break; }   /* spline_manager.nodes[node_idx] :530 -> IndexError */
            const double* a = na + (size_t)node_idx * NA;
            if (a[A_TURN] != 0) {   /* handle_turn :487-507 */
                double angle = a[A_TURN] * (PI / 180.0);    /* np.radians */
                long K = ora_motion_profile_angle(angle, V, max_acc, w, dt, tcap, ins_h, ins_w);
                if (K < 0) { rc = ORA_ERR_CAPACITY; break; }
                if (o.T == 0) { rc = ORA_ERR_INDEX; break; }
                double start_heading = o.head[o.T - 1];
                double lp = o.pos[o.T - 1], lx = o.x[o.T - 1], ly = o.y[o.T - 1];
                int ok = 1;
                for (long i = 0; i < K && ok; i++) {
                    while (ins_h[i] + start_heading > PI) ins_h[i] -= 2 * PI;
                    while (ins_h[i] + start_heading < -PI) ins_h[i] += 2 * PI;
                }
                for (long i = 0; i < K && ok; i++)
                    ok = push(&o, current_time + (double)i * dt, lp, 0.0, 0.0, start_heading + ins_h[i], ins_w[i], lx, ly);
                if (!ok) { rc = ORA_ERR_CAPACITY; break; }
                current_time = current_time + (double)K * dt;
            }
            if (nf[node_idx] & F_REVERSE) is_reversed = !is_reversed;
            if (a[A_WAIT] > 0) {
                if (o.T == 0) { rc = ORA_ERR_INDEX; break; }
                if (!do_wait(&o, a[A_WAIT], &current_time, dt)) { rc = ORA_ERR_CAPACITY; break; }
            }
        }
        if (action_idx < A) {
            const double* p = apa + (size_t)action_idx * APA;
            if (prev_t < p[P_T] && p[P_T] < t) {
                actions_map[am++] = o.T;
                if (p[P_WAIT] > 0) {
                    if (o.T == 0) { rc = ORA_ERR_INDEX; break; }
                    if (!do_wait(&o, p[P_WAIT], &current_time, dt)) { rc = ORA_ERR_CAPACITY; break; }
                }
                action_idx += 1;
            }
        }
        prev_t = t;
        double curvature = ora_snap(P, n, kap, t);
        double heading = ora_snap(P, n, th, t) - (is_reversed ? PI : 0);
        heading = pymod(heading + PI, 2 * PI) - PI;
        heading *= -1;
        double c[2];
        ora_eval(S, first_node, param_end, seg, 0, t, c);
        double tv = ora_lerp_uniform(current_pos, dd, D, vel);
        double ntv = ora_lerp_uniform(current_pos + dd, dd, D, vel);
        tv = pymax((tv + ntv) / 2, 0.001);
        double accel = (tv - current_vel) / dt;
        accel = fmin(fmax(accel, -max_dec), max_acc);           /* np.clip */
        double angular_vel = tv * curvature * -1;
        current_vel = fmin(fmax(current_vel + accel * dt, 0.0), tv);
        double dpos = current_vel * dt + 0.5 * accel * dt * dt;
        if (current_vel <= 0.1) dpos = 0.1 * dt + 0.5 * accel * dt * dt;
        current_pos += dpos;
        double sgn = is_reversed ? -1.0 : 1.0;
        if (!push(&o, current_time, current_pos, current_vel * sgn, accel * sgn, heading, angular_vel, c[0], c[1])) {
            rc = ORA_ERR_CAPACITY; break;
        }
        current_time += dt;
    }
    free(ins_h); free(ins_w);
    if (rc < 0) return rc;
    nodes_map[nm++] = o.T;
    *n_nm = nm; *n_am = am;
    return o.T;
}

/* ------------------------------------------------------------------------------------------
 * S1'. Gauss-Legendre arc length and bisection inverse (quintic_hermite_spline.py:592-717).
 * pts / wts = np.polynomial.legendre.leggauss(npts) evaluated on the host.
 * ------------------------------------------------------------------------------------------ */
static double np_sum(const double* a, int n)
{   /* numpy pairwise_sum for n <= 128 (A.9) */
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r += a[i]; return r; }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

int ora_gl_arclen(int nseg, double pend, const double* seg, double t0, double t1, int npts,
                  const double* pts, const double* wts, double* out)
{
    if (t0 >= t1) return ORA_ERR_VALUE;
    if (t0 < 0.0 || t1 > pend) return ORA_ERR_VALUE;
    if (npts > 128) return ORA_ERR_CAPACITY;
    double half = (t1 - t0) / 2, mid = (t0 + t1) / 2;
    double wm[128];
    for (int j = 0; j < npts; j++) {
        double tau = pts[j] * half + mid;
        double d[2];
        ora_eval_spline(nseg, pend, seg, 1, tau, d);
        wm[j] = wts[j] * norm1d(d[0], d[1]);
    }
    *out = half * np_sum(wm, npts);
    return 0;
}

int ora_gl_inverse(int nseg, double pend, const double* seg, double s, double tol, int max_iter, int npts,
                   const double* pts, const double* wts, double* out)
{
    if (s < 0) return ORA_ERR_VALUE;
    double total;
    int r = ora_gl_arclen(nseg, pend, seg, 0.0, pend, npts, pts, wts, &total);
    if (r < 0) return r;
    if (s > total) return ORA_ERR_VALUE;
    if (s == 0) { *out = 0.0; return 0; }
    if (s == total) { *out = pend; return 0; }
    double lo = 0.0, hi = pend;
    for (int it = 0; it < max_iter; it++) {
        double m = (lo + hi) / 2, len;
        r = ora_gl_arclen(nseg, pend, seg, 0.0, m, npts, pts, wts, &len);
        if (r < 0) return r;
        double err = len - s;
        if (fabs(err) < tol) { *out = m; return 0; }
        if (err > 0) hi = m; else lo = m;
    }
    *out = (lo + hi) / 2;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Whole path in one call (build_path + generate_motion_profile), used for batch parity checks and as
 * the timed CPU baseline.  Scratch is malloc'ed per call.  Returns T or a negative error;
 * summary[5] = {T, total_length, t_end, max|v|, status}.
 * ------------------------------------------------------------------------------------------ */
long ora_full(int n, const double* na, const int* nf, int A, const double* apa, const int* apf,
              const double* cons, double dt, double dd, long cap_d, long cap_t, double* vel_out, long* D_out,
              double* times, double* pos, double* lin, double* acc, double* head, double* ang, double* xo,
              double* yo, long* nodes_map, int* n_nm, long* actions_map, int* n_am, double* summary)
{
    long rc = 0;
    int G = n - 1;
    double* seg = (double*)malloc(sizeof(double) * 12 * (size_t)(G > 0 ? G : 1));
    int* fn = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    double* pe = (double*)malloc(sizeof(double) * (size_t)n);
    double* sl = (double*)malloc(sizeof(double) * (size_t)n);
    double* pc = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double *lut_d = NULL, *lut_t = NULL, *kap = NULL, *th = NULL, *tq = NULL, *kq = NULL, *hq = NULL;
    double* ma = (double*)malloc(sizeof(double) * (size_t)(n + A + 3));
    long* bidx = (long*)malloc(sizeof(long) * (size_t)(n + A + 3));
    int* bval = (int*)malloc(sizeof(int) * (size_t)(n + A + 3));
    long D = 0, T = 0;
    double total = 0.0;
    int S = ora_build_path(n, na, nf, seg, fn, pe, sl, pc);
    if (S < 0) { rc = S; goto done; }
    long Q = 1000L * S, P = 1000L * n;
    lut_d = (double*)malloc(sizeof(double) * (size_t)Q);
    lut_t = (double*)malloc(sizeof(double) * (size_t)Q);
    kap = (double*)malloc(sizeof(double) * (size_t)P);
    th = (double*)malloc(sizeof(double) * (size_t)P);
    total = ora_build_lut(S, fn, pe, seg, 1000, lut_d, lut_t);
    ora_build_props(n, S, fn, pe, seg, 1000, kap, th);
    tq = (double*)malloc(sizeof(double) * (size_t)cap_d);
    kq = (double*)malloc(sizeof(double) * (size_t)cap_d);
    hq = (double*)malloc(sizeof(double) * (size_t)cap_d);
    int n_acc = 0, n_b = 0;
    D = ora_dist_sample(n, na, nf, A, apa, apf, cons, dd, 0.01, Q, lut_d, lut_t, total, P, kap, th, cap_d, tq,
                        kq, hq, vel_out, ma, &n_acc, bidx, bval, &n_b);
    if (D < 0) { rc = D; goto done; }
    ora_fwd_bwd(D, kq, hq, vel_out, cons, dd, ma, n_b, bidx, bval, 0.01, 0.01, 0);
    T = ora_profile(n, na, nf, A, apa, apf, cons, dt, dd, S, fn, pe, seg, Q, lut_d, lut_t, total, P, kap, th, D,
                    vel_out, cap_t, times, pos, lin, acc, head, ang, xo, yo, nodes_map, n_nm, actions_map, n_am);
    if (T < 0) { rc = T; goto done; }
    rc = T;
done:
    if (D_out) *D_out = D;
    if (summary) {
        double mv = 0.0;
        for (long i = 0; i < T; i++) if (fabs(lin[i]) > mv) mv = fabs(lin[i]);
        summary[0] = (double)(T > 0 ? T : 0);
        summary[1] = total;
        summary[2] = T > 0 ? times[T - 1] : 0.0;
        summary[3] = mv;
        summary[4] = rc < 0 ? (double)rc : 0.0;
    }
    free(seg); free(fn); free(pe); free(sl); free(pc); free(lut_d); free(lut_t); free(kap); free(th);
    free(tq); free(kq); free(hq); free(ma); free(bidx); free(bval);
    return rc;
}

/* number of OpenMP threads ora_full_batch uses (0 = library default) */
void ora_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Batch driver for timing: B paths with identical node count, OpenMP over paths; only summaries are
 * kept (outputs go to per-thread scratch).  node_attr[B][n][12] etc. */
void ora_full_batch(long B, int n, const double* na, const int* nf, int Amax, const int* n_ap, const double* apa,
                    const int* apf, const double* cons, double dt, double dd, long cap_d, long cap_t,
                    double* summaries)
{
#pragma omp parallel
    {
        double* buf = (double*)malloc(sizeof(double) * (size_t)(cap_d + 8 * cap_t));
        long* nmap = (long*)malloc(sizeof(long) * (size_t)(n + 2));
        long* amap = (long*)malloc(sizeof(long) * (size_t)(Amax + 2));
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < B; b++) {
            double* v = buf;
            double* o = buf + cap_d;
            long D;
            int nnm, nam;
            ora_full(n, na + (size_t)b * n * NA, nf + (size_t)b * n, n_ap ? n_ap[b] : 0,
                     apa ? apa + (size_t)b * Amax * APA : NULL, apf ? apf + (size_t)b * Amax : NULL,
                     cons + (size_t)b * 6, dt, dd, cap_d, cap_t, v, &D, o, o + cap_t, o + 2 * cap_t, o + 3 * cap_t,
                     o + 4 * cap_t, o + 5 * cap_t, o + 6 * cap_t, o + 7 * cap_t, nmap, &nnm, amap, &nam,
                     summaries + (size_t)b * 5);
        }
        free(buf); free(nmap); free(amap);
    }
}

/* Like ora_full_batch, but also keeps every INTEGER output of each path (tests compare them between arithmetic
 * modes / against the engine at scale): ints[b][0..2] = D, len(nodes_map), len(actions_map), then nodes_map
 * (n + 1 slots) and actions_map (Amax slots); K = 3 + (n + 1) + Amax longs per path.  checks[b][0..3] = sums of
 * the velocity row, of times, of x and of y (cheap value fingerprints). */
void ora_full_batch_ex(long B, int n, const double* na, const int* nf, int Amax, const int* n_ap, const double* apa,
                       const int* apf, const double* cons, double dt, double dd, long cap_d, long cap_t,
                       double* summaries, long* ints, double* checks)
{
    const long K = 3 + (n + 1) + Amax;
#pragma omp parallel
    {
        double* buf = (double*)malloc(sizeof(double) * (size_t)(cap_d + 8 * cap_t));
        long* nmap = (long*)malloc(sizeof(long) * (size_t)(n + 2));
        long* amap = (long*)malloc(sizeof(long) * (size_t)(Amax + 2));
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < B; b++) {
            double* v = buf;
            double* o = buf + cap_d;
            long D = 0;
            int nnm = 0, nam = 0;
            long T = ora_full(n, na + (size_t)b * n * NA, nf + (size_t)b * n, n_ap ? n_ap[b] : 0,
                              apa ? apa + (size_t)b * Amax * APA : NULL, apf ? apf + (size_t)b * Amax : NULL,
                              cons + (size_t)b * 6, dt, dd, cap_d, cap_t, v, &D, o, o + cap_t, o + 2 * cap_t,
                              o + 3 * cap_t, o + 4 * cap_t, o + 5 * cap_t, o + 6 * cap_t, o + 7 * cap_t, nmap, &nnm,
                              amap, &nam, summaries + (size_t)b * 5);
            long* r = ints + (size_t)b * K;
            for (long q = 0; q < K; q++) r[q] = -1;
            r[0] = D; r[1] = nnm; r[2] = nam;
            for (int q = 0; q < nnm && q < n + 1; q++) r[3 + q] = nmap[q];
            for (int q = 0; q < nam && q < Amax; q++) r[3 + (n + 1) + q] = amap[q];
            double* c = checks + (size_t)b * 4;
            c[0] = c[1] = c[2] = c[3] = 0.0;
            if (T > 0) {
                for (long i = 0; i < D; i++) c[0] += v[i];
                for (long i = 0; i < T; i++) { c[1] += o[i]; c[2] += o[6 * cap_t + i]; c[3] += o[7 * cap_t + i]; }
            }
        }
        free(buf); free(nmap); free(amap);
    }
}
