/*
 * cmt_oracle.c -- CPU restatement of the reference's Monte Carlo propagation path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * shipped path (libcmt_b200.so, CUDA) never links, imports or calls it.
 *
 * What it restates (paths relative to /root/reference/src/trajectories):
 *   molecule.py:26-68                      Molecule.x/v/update_trajectory
 *   beamline.py:20-38                      Beamline.propagate_through
 *   beamline_elements/apertures.py:92-115  CircularAperture.propagate_through
 *   beamline_elements/apertures.py:157-189 RectangularAperture
 *   beamline_elements/apertures.py:221-270 FieldPlates
 *   beamline_elements/electrostatic_lens.py:48-118,215-228  ElectrostaticLens
 *   beamline_elements/meshes.py:84-117     Honeycomb.propagate_through (see below)
 *   distributions.py:69-76,112-119,155-162 (distribution shapes; the bit stream
 *                                           is the build's Philox, not NumPy's)
 *   trajectory_simulator.py:62-76          per-molecule loop + Counter
 *
 * Parity status: PINNED by execution.  The reference ships no tests or golden
 * vectors, so tests/golden/make_golden.py runs the unmodified reference source
 * in the build container on fixed initial conditions and commits the results;
 * tests/test_oracle_golden.py requires this file to reproduce them bit for bit
 * (same numpy 2.3.5 / scipy 1.18.1 / glibc as the container).  The Stark-curve
 * producer (centrex_TlF, external and unpinned) is NOT restated here: the lens
 * acceleration table is an input ("parity unpinned" at that boundary only).
 *
 * Honeycomb: the reference's own logic (stepping, nearest-centre argmin, `if not idx`,
 * fate) is pinned by tests/golden/honeycomb.npz, but its geometry comes from two
 * third-party packages that are absent here and pinned nowhere by the reference:
 * hexalattice.make_grid (cell centres) and matplotlib's RegularPolygon.contains_point
 * (hit test).  Both are restated from their published algorithms -- PARITY UNPINNED at
 * that boundary: the golden file was produced with the restatements in oracle/stubs
 * standing in for the real packages.
 *
 * Arithmetic notes that matter for bit parity with the reference as executed:
 *   - `delta_t**2` on a numpy float64 *scalar* goes through libm pow(), which
 *     is not always equal to dt*dt (0.08 % of random inputs differ by 1 ulp);
 *     we call pow() too.  `x[:2]**2` on an *array* is an exact square.
 *   - scipy 1.18.1 interp1d(kind="linear") evaluates through np.interp:
 *     exact table value when r == x_j, else slope*(r-x_j)+y_j with
 *     slope=(y_{j+1}-y_j)/(x_{j+1}-x_j) computed on the fly, no FMA.
 *   - Build with -ffp-contract=off; no -ffast-math; no -march flags.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_CIRCULAR 0
#define ORC_RECTANGULAR 1
#define ORC_FIELDPLATES 2
#define ORC_LENS 3
#define ORC_HONEYCOMB 4

typedef struct {
    int32_t type;     /* ORC_* */
    int32_t fate;     /* fate id on hit; lens: id of "Lens entrance" */
    int32_t fate2;    /* lens: id of "Inside lens" */
    int32_t table;    /* lens: table index */
    int32_t n_steps;  /* lens: int(rint(L/dz)), electrostatic_lens.py:87; honeycomb: nx */
    int32_t pad_;     /* honeycomb: ny */
    double z0, z1;
    double x1, x2, y1, y2; /* rectangular / field-plate edges, apertures.py:157-163,221-225 */
    double R;              /* d/2 for circular aperture and lens bore; honeycomb: polygon radius, meshes.py:73-77 */
    double dz;             /* lens step, electrostatic_lens.py:30; honeycomb: pitch (min_diam), meshes.py:58;
                              honeycomb also: x1 = mid_x, y1 = mid_y of make_grid */
} orc_element;

typedef struct {
    const orc_element *el;
    int n_el;
    const double *tab_r;   /* concatenated table abscissae */
    const double *tab_a;   /* concatenated table ordinates  */
    const int32_t *tab_off;
    const int32_t *tab_len;
    int fate_detected;
    double g;
} orc_beamline;

/* One molecule = the last trajectory row (molecule.py:26-56 read row n-1). */
typedef struct {
    double x[3], v[3], a[3], t;
    int alive;
    int fate;
    int n_rows;
    double *rows;      /* optional (max_rows,10): x,y,z,vx,vy,vz,ax,ay,az,t */
    long max_rows;
    int64_t planes, steps, oob;
} orc_mol;

static void store_row(orc_mol *m)
{
    /* Trajectory.update, molecule.py:133-144 */
    if (m->rows && m->n_rows < m->max_rows) {
        double *r = m->rows + (long)m->n_rows * 10;
        r[0] = m->x[0]; r[1] = m->x[1]; r[2] = m->x[2];
        r[3] = m->v[0]; r[4] = m->v[1]; r[5] = m->v[2];
        r[6] = m->a[0]; r[7] = m->a[1]; r[8] = m->a[2];
        r[9] = m->t;
    }
    m->n_rows++;
}

/*
 * gcc folds pow(x, 2.0) into x*x even without -ffast-math; NumPy's scalar power
 * really calls libm, so go through a volatile pointer to keep the call.
 */
static double (*volatile libm_pow)(double, double) = pow;

/* Molecule.x(delta_t), molecule.py:26-34: `if not delta_t` returns the row. */
static void pos_after(const orc_mol *m, double dt, double out[3])
{
    if (dt == 0.0) { out[0] = m->x[0]; out[1] = m->x[1]; out[2] = m->x[2]; return; }
    double dt2 = libm_pow(dt, 2.0);         /* numpy scalar ** 2 -> libm pow */
    for (int c = 0; c < 3; c++)
        out[c] = (m->x[c] + m->v[c] * dt) + (m->a[c] * dt2) / 2;
}

/* Molecule.v(delta_t), molecule.py:36-44 */
static void vel_after(const orc_mol *m, double dt, double out[3])
{
    if (dt == 0.0) { out[0] = m->v[0]; out[1] = m->v[1]; out[2] = m->v[2]; return; }
    for (int c = 0; c < 3; c++) out[c] = m->v[c] + m->a[c] * dt;
}

/* Molecule.update_trajectory, molecule.py:58-68: new row stores default a. */
static void update_trajectory(orc_mol *m, double dt, double g)
{
    double xn[3], vn[3];
    pos_after(m, dt, xn);
    vel_after(m, dt, vn);
    double tn = m->t + dt;
    for (int c = 0; c < 3; c++) { m->x[c] = xn[c]; m->v[c] = vn[c]; }
    m->a[0] = 0.0; m->a[1] = -g; m->a[2] = 0.0;
    m->t = tn;
    m->planes++;
    store_row(m);
}

static void mark_dead(orc_mol *m, int fate) { m->alive = 0; m->fate = fate; }

static double rho_of(const double x[3])
{
    /* np.sqrt(np.sum(x[:2]**2)), apertures.py:110 */
    return sqrt(x[0] * x[0] + x[1] * x[1]);
}

/* CircularAperture.propagate_through, apertures.py:92-115 */
static void circular(const orc_element *e, orc_mol *m, double g)
{
    const double zs[2] = { e->z0, e->z1 };
    for (int k = 0; k < 2; k++) {
        double dt = (zs[k] - m->x[2]) / m->v[2];
        update_trajectory(m, dt, g);
        if (rho_of(m->x) > e->R) { mark_dead(m, e->fate); return; }
    }
}

/* RectangularAperture.propagate_through, apertures.py:165-189 */
static void rectangular(const orc_element *e, orc_mol *m, double g)
{
    const double zs[2] = { e->z0, e->z1 };
    for (int k = 0; k < 2; k++) {
        double dt = (zs[k] - m->x[2]) / m->v[2];
        update_trajectory(m, dt, g);
        int inside = (e->x1 < m->x[0] && m->x[0] < e->x2) &&
                     (e->y1 < m->x[1] && m->x[1] < e->y2);
        if (!inside) { mark_dead(m, e->fate); return; }
    }
}

/* FieldPlates.propagate_through, apertures.py:227-270 */
static void fieldplates(const orc_element *e, orc_mol *m, double g)
{
    double dt = (e->z0 - m->x[2]) / m->v[2];
    update_trajectory(m, dt, g);
    if (!(e->x1 < m->x[0] && m->x[0] < e->x2)) { mark_dead(m, e->fate); return; }

    dt = (e->z1 - m->x[2]) / m->v[2];
    double xn[3];
    pos_after(m, dt, xn);
    if (!(e->x1 < xn[0] && xn[0] < e->x2)) {
        if (m->v[0] < 0) dt = (e->x1 - m->x[0]) / m->v[0];
        else if (m->v[0] > 0) dt = (e->x2 - m->x[0]) / m->v[0];
        update_trajectory(m, dt, g);
        mark_dead(m, e->fate);
        return;
    }
    update_trajectory(m, dt, g);
}

/*
 * a_interp(r): scipy interp1d(kind="linear") -> np.interp on a sorted table.
 * Out-of-range r makes the reference raise ValueError and abort the whole run
 * (bounds_error=True).  The build defines that case instead: evaluate on the
 * nearest end interval's line and count it in `oob`; the Python layer turns a
 * non-zero count into the same ValueError.
 */
static double table_eval(const double *xp, const double *fp, int n, double x, int64_t *oob)
{
    if (isnan(x)) return x;
    int j;
    if (x > xp[n - 1]) { (*oob)++; j = n - 2; }
    else if (x < xp[0]) { (*oob)++; j = 0; }
    else {
        int lo = 0, hi = n;             /* largest j with xp[j] <= x */
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (x >= xp[mid]) lo = mid; else hi = mid; }
        j = lo;
        if (j == n - 1) return fp[j];
        if (xp[j] == x) return fp[j];
    }
    double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double res = slope * (x - xp[j]) + fp[j];
    if (isnan(res)) {
        res = slope * (x - xp[j + 1]) + fp[j + 1];
        if (isnan(res) && fp[j] == fp[j + 1]) res = fp[j];
    }
    return res;
}

/* ElectrostaticLens.lens_acceleration, electrostatic_lens.py:215-228 */
static void lens_acc(const orc_beamline *b, const orc_element *e, const double x[3],
                     double a[3], int64_t *oob)
{
    const double *xp = b->tab_r + b->tab_off[e->table];
    const double *fp = b->tab_a + b->tab_off[e->table];
    int n = b->tab_len[e->table];
    double r = sqrt(x[0] * x[0] + x[1] * x[1]);
    double a_r = table_eval(xp, fp, n, r, oob);
    a[0] = 0.0; a[1] = 0.0; a[2] = 0.0;
    if (r != 0) {
        a[0] = a_r * x[0] / r;
        a[1] = a_r * x[1] / r;
        a[2] = 0;
    }
    a[1] -= b->g;
}

/* ElectrostaticLens.propagate_through + propagate_inside_lens, electrostatic_lens.py:48-118 */
static void lens(const orc_beamline *b, const orc_element *e, orc_mol *m)
{
    double dt = (e->z0 - m->x[2]) / m->v[2];
    update_trajectory(m, dt, b->g);
    if (rho_of(m->x) > e->R) { mark_dead(m, e->fate); return; }

    const int N = e->n_steps;
    dt = e->dz / m->v[2];
    for (int i = 0; i < N; i++) {
        double x[3], k1[3], k2[3], k3[3], k4[3], l1[3], l2[3], l3[3], l4[3], p[3];
        for (int c = 0; c < 3; c++) { x[c] = m->x[c]; k1[c] = m->v[c]; }
        lens_acc(b, e, x, l1, &m->oob);

        for (int c = 0; c < 3; c++) k2[c] = k1[c] + dt * l1[c] / 2;
        for (int c = 0; c < 3; c++) p[c] = x[c] + dt * k1[c];
        lens_acc(b, e, p, l2, &m->oob);

        for (int c = 0; c < 3; c++) k3[c] = k1[c] + dt * l2[c] / 2;
        for (int c = 0; c < 3; c++) p[c] = x[c] + dt * k2[c] / 2;
        lens_acc(b, e, p, l3, &m->oob);

        for (int c = 0; c < 3; c++) k4[c] = k1[c] + dt * l3[c];
        for (int c = 0; c < 3; c++) p[c] = x[c] + dt * k3[c];
        lens_acc(b, e, p, l4, &m->oob);

        for (int c = 0; c < 3; c++) {
            m->x[c] = x[c] + dt * (k1[c] + 2 * k2[c] + 2 * k3[c] + k4[c]) / 6;
            m->v[c] = k1[c] + dt * (l1[c] + 2 * l2[c] + 2 * l3[c] + l4[c]) / 6;
            m->a[c] = l1[c];
        }
        m->t = m->t + dt;
        m->steps++;
        store_row(m);

        if (rho_of(m->x) > e->R) { mark_dead(m, e->fate2); return; }
    }

    /* exit step uses the last stored a (= l1 of the final RK step), molecule.py:46-50 */
    dt = (e->z1 - m->x[2]) / m->v[2];
    update_trajectory(m, dt, b->g);
}

/* ---- Honeycomb, meshes.py:84-117 ------------------------------------------ */
/* np.cos / np.sin of 2*pi/6*k + pi/2 (Path.unit_regular_polygon(6)); checked against NumPy by the tests */
static const double hex_ux[6] = {0x1.1a62633145c07p-54, -0x1.bb67ae8584ca9p-1, -0x1.bb67ae8584cacp-1,
                                 -0x1.a79394c9e8a0ap-53, 0x1.bb67ae8584ca8p-1, 0x1.bb67ae8584caep-1};
static const double hex_uy[6] = {0x1.0000000000000p+0, 0x1.0000000000003p-1, -0x1.ffffffffffffbp-2,
                                 -0x1.0000000000000p+0, -0x1.0000000000004p-1, 0x1.ffffffffffff3p-2};
#define HEX_RATIO 0x1.bb67ae8584caap-1 /* np.sqrt(3) / 2 */

static void hex_centre(const orc_element *e, long idx, double *xc, double *yc)
{
    /* hexalattice.make_grid: row-major, odd rows shifted by half a pitch, middle cell moved to the origin */
    long row = idx / e->n_steps, col = idx % e->n_steps;
    double cx = (double)col;
    if (row & 1) cx += 0.5;
    *xc = cx * e->dz - e->x1;
    *yc = ((double)row * HEX_RATIO) * e->dz - e->y1;
}

static long hex_nearest(const orc_element *e, double px, double py)
{
    /* idx = np.argmin(np.sqrt((px - xcoords)**2 + (py - ycoords)**2)), meshes.py:105-109: every centre, first minimum;
     * a NaN distance wins over everything that follows (np.argmin propagates the first NaN) */
    long n = (long)e->n_steps * e->pad_, best = 0;
    double best_rho = 0.0;
    for (long i = 0; i < n; i++) {
        double xc, yc;
        hex_centre(e, i, &xc, &yc);
        double dx = px - xc, dy = py - yc;
        double rho = sqrt(dx * dx + dy * dy);
        if (rho != rho) return i;
        if (i == 0 || rho < best_rho) { best_rho = rho; best = i; }
    }
    return best;
}

static int hex_contains(const orc_element *e, long idx, double tx, double ty)
{
    /* RegularPolygon((xc, yc), 6, radius=R).contains_point((tx, ty)): crossings-multiply test of matplotlib's
     * _path.h on the vertices unit * R + centre; non-finite points are outside */
    if (!isfinite(tx) || !isfinite(ty)) return 0;
    double xc, yc;
    hex_centre(e, idx, &xc, &yc);
    int inside = 0;
    for (int k = 0; k < 6; k++) {
        int k1 = (k + 1) % 6;
        double x0 = hex_ux[k] * e->R + xc, y0 = hex_uy[k] * e->R + yc;
        double x1 = hex_ux[k1] * e->R + xc, y1 = hex_uy[k1] * e->R + yc;
        int f0 = y0 >= ty, f1 = y1 >= ty;
        if (f0 != f1 && (((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1)) inside ^= 1;
    }
    return inside;
}

static void honeycomb(const orc_element *e, orc_mol *m, double g)
{
    long idx = 0;                                   /* `idx = None`, and `if not idx` is also true for cell 0 */
    const double zs[2] = {e->z0, e->z1};
    for (int k = 0; k < 2; k++) {
        double dt = (zs[k] - m->x[2]) / m->v[2];
        update_trajectory(m, dt, g);
        if (idx == 0) idx = hex_nearest(e, m->x[0], m->x[1]);
        if (!hex_contains(e, idx, m->x[0], m->x[1])) { mark_dead(m, e->fate); return; }
    }
}

static void propagate_one(const orc_beamline *b, orc_mol *m)
{
    /* Beamline.propagate_through, beamline.py:20-38 */
    for (int i = 0; i < b->n_el; i++) {
        const orc_element *e = &b->el[i];
        switch (e->type) {
        case ORC_CIRCULAR:    circular(e, m, b->g); break;
        case ORC_RECTANGULAR: rectangular(e, m, b->g); break;
        case ORC_FIELDPLATES: fieldplates(e, m, b->g); break;
        case ORC_LENS:        lens(b, e, m); break;
        case ORC_HONEYCOMB:   honeycomb(e, m, b->g); break;
        default: break;
        }
        if (!m->alive) break;
    }
    if (m->alive) m->fate = b->fate_detected;
}

static void init_mol(orc_mol *m, const double x0[3], const double v0[3], double g,
                     double *rows, long max_rows)
{
    /* Molecule.init_trajectory, molecule.py:15-24: row 0 = (x0, v0, (0,-g,0), 0) */
    memset(m, 0, sizeof(*m));
    for (int c = 0; c < 3; c++) { m->x[c] = x0[c]; m->v[c] = v0[c]; }
    m->a[0] = 0.0; m->a[1] = -g; m->a[2] = 0.0;
    m->t = 0.0;
    m->alive = 1;
    m->fate = -1;
    m->rows = rows;
    m->max_rows = max_rows;
    store_row(m);
}

/*
 * Propagate n molecules from explicit initial conditions.
 *   ic       (6,n) row-major: x,y,z,vx,vy,vz
 *   fate     (n) int32
 *   fin      (10,n) row-major or NULL: x,y,z,vx,vy,vz,ax,ay,az,t of the last row
 *   n_rows   (n) int32 or NULL
 *   rows     (n,max_rows,10) or NULL
 *   counters (n_fates) int64, accumulated
 *   work     (3) int64, accumulated: ballistic rows, lens RK steps, table out-of-range evals
 */
int orc_propagate(const orc_element *el, int n_el, const double *tab_r, const double *tab_a,
                  const int32_t *tab_off, const int32_t *tab_len, int fate_detected, double g,
                  long n, const double *ic, int32_t *fate, double *fin, int32_t *n_rows,
                  double *rows, long max_rows, int64_t *counters, int n_fates, int64_t *work,
                  int n_threads)
{
    orc_beamline b = { el, n_el, tab_r, tab_a, tab_off, tab_len, fate_detected, g };
    int64_t w0 = 0, w1 = 0, w2 = 0;
    int64_t *cnt = counters;
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel reduction(+ : w0, w1, w2)
#endif
    {
        int64_t *lc = (int64_t *)calloc((size_t)(n_fates > 0 ? n_fates : 1), sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
        for (long i = 0; i < n; i++) {
            double x0[3] = { ic[0 * n + i], ic[1 * n + i], ic[2 * n + i] };
            double v0[3] = { ic[3 * n + i], ic[4 * n + i], ic[5 * n + i] };
            orc_mol m;
            init_mol(&m, x0, v0, g, rows ? rows + i * max_rows * 10 : NULL, max_rows);
            propagate_one(&b, &m);
            if (fate) fate[i] = m.fate;
            if (n_rows) n_rows[i] = m.n_rows;
            if (fin) {
                for (int c = 0; c < 3; c++) {
                    fin[(0 + c) * n + i] = m.x[c];
                    fin[(3 + c) * n + i] = m.v[c];
                    fin[(6 + c) * n + i] = m.a[c];
                }
                fin[9 * n + i] = m.t;
            }
            if (cnt && m.fate >= 0 && m.fate < n_fates) lc[m.fate]++;
            w0 += m.planes; w1 += m.steps; w2 += m.oob;
        }
        if (cnt) {
#ifdef _OPENMP
#pragma omp critical
#endif
            for (int k = 0; k < n_fates; k++) cnt[k] += lc[k];
        }
        free(lc);
    }
    if (work) { work[0] += w0; work[1] += w1; work[2] += w2; }
    return 0;
}

/* ------------------------------------------------------------------------- *
 * Counter-based source: Philox4x32-10 (Salmon et al., SC'11), keyed by the run
 * seed and indexed by the GLOBAL molecule index, so results do not depend on
 * how molecules are split over threads, chunks or GPUs.
 *
 *   counter = (index_lo, index_hi, block, 0), key = (seed_lo, seed_hi)
 *   block 0 -> (u0,u1): Box-Muller -> (nx, ny)   velocity x,y
 *   block 1 -> (u0,u1): Box-Muller -> (nz, ns)   velocity z, spare normal
 *   block 2 -> (u0,u1): position uniforms / second position normal pair
 *   u = ((w >> 11) + 0.5) * 2^-53 with w = hi32:lo32 of two output words
 *
 *   kind 0 (distributions.py:112-119, CeNTREXPositionDistribution):
 *       theta = u0*2*pi; r = sqrt(u1)*d/2; x = r*cos(theta); y = r*sin(theta)
 *   kind 1 (distributions.py:155-162, GaussianPositionDistribution):
 *       (n0,n1) = Box-Muller(block 2); x = 0 + sigmax*n0; y = 0 + sigmay*n1
 *   velocity (distributions.py:69-76): v_c = mean_c + sigma_c * n_c
 * ------------------------------------------------------------------------- */
typedef struct {
    int32_t pos_kind; /* 0 = uniform disc, 1 = gaussian */
    int32_t pad_;
    double vmean[3], vsigma[3];
    double p0, p1;    /* disc: p0 = d/2 ; gaussian: sigmax, sigmay */
    double z;
} orc_source;

static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k[2] = { key[0], key[1] };
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

static inline void uniforms(uint64_t seed, uint64_t index, uint32_t block, double *u0, double *u1)
{
    uint32_t ctr[4] = { (uint32_t)index, (uint32_t)(index >> 32), block, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    uint64_t w0 = ((uint64_t)o[1] << 32) | o[0];
    uint64_t w1 = ((uint64_t)o[3] << 32) | o[2];
    *u0 = ((double)(w0 >> 11) + 0.5) * 0x1.0p-53;
    *u1 = ((double)(w1 >> 11) + 0.5) * 0x1.0p-53;
}

static inline void box_muller(double u0, double u1, double *n0, double *n1)
{
    double rad = sqrt(-2.0 * log(u0));
    double ang = 6.283185307179586476925286766559 * u1;
    *n0 = rad * cos(ang);
    *n1 = rad * sin(ang);
}

static void draw_one(const orc_source *s, uint64_t seed, uint64_t index, double x0[3], double v0[3])
{
    double u0, u1, n0, n1, n2, n3;
    uniforms(seed, index, 0, &u0, &u1);
    box_muller(u0, u1, &n0, &n1);
    uniforms(seed, index, 1, &u0, &u1);
    box_muller(u0, u1, &n2, &n3);
    v0[0] = s->vmean[0] + s->vsigma[0] * n0;
    v0[1] = s->vmean[1] + s->vsigma[1] * n1;
    v0[2] = s->vmean[2] + s->vsigma[2] * n2;
    uniforms(seed, index, 2, &u0, &u1);
    if (s->pos_kind == 0) {
        double theta = 6.283185307179586476925286766559 * u0;
        double r = sqrt(u1) * s->p0;
        x0[0] = r * cos(theta);
        x0[1] = r * sin(theta);
    } else {
        box_muller(u0, u1, &n0, &n1);
        x0[0] = s->p0 * n0;
        x0[1] = s->p1 * n1;
    }
    x0[2] = s->z;
    (void)n3;
}

/* Fill ic (6,n) for global indices [first, first+n). */
int orc_draw(const orc_source *s, uint64_t seed, uint64_t first, long n, double *ic, int n_threads)
{
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static)
#endif
    for (long i = 0; i < n; i++) {
        double x0[3], v0[3];
        draw_one(s, seed, first + (uint64_t)i, x0, v0);
        for (int c = 0; c < 3; c++) { ic[c * n + i] = x0[c]; ic[(3 + c) * n + i] = v0[c]; }
    }
    return 0;
}

/*
 * Whole run on the host cores: draw + propagate + count, nothing stored per
 * molecule.  This is what bench.py times as the CPU baseline (kind "port").
 */
int orc_run(const orc_element *el, int n_el, const double *tab_r, const double *tab_a,
            const int32_t *tab_off, const int32_t *tab_len, int fate_detected, double g,
            const orc_source *src, uint64_t seed, uint64_t first, long n,
            int64_t *counters, int n_fates, int64_t *work, int n_threads)
{
    orc_beamline b = { el, n_el, tab_r, tab_a, tab_off, tab_len, fate_detected, g };
    int64_t w0 = 0, w1 = 0, w2 = 0;
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel reduction(+ : w0, w1, w2)
#endif
    {
        int64_t *lc = (int64_t *)calloc((size_t)(n_fates > 0 ? n_fates : 1), sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
        for (long i = 0; i < n; i++) {
            double x0[3], v0[3];
            draw_one(src, seed, first + (uint64_t)i, x0, v0);
            orc_mol m;
            init_mol(&m, x0, v0, g, NULL, 0);
            propagate_one(&b, &m);
            if (m.fate >= 0 && m.fate < n_fates) lc[m.fate]++;
            w0 += m.planes; w1 += m.steps; w2 += m.oob;
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int k = 0; k < n_fates; k++) counters[k] += lc[k];
        free(lc);
    }
    if (work) { work[0] += w0; work[1] += w1; work[2] += w2; }
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_sizeof_element(void) { return (int)sizeof(orc_element); }
int orc_sizeof_source(void) { return (int)sizeof(orc_source); }
