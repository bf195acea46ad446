/*
 * cmt.h -- C ABI of libcmt_b200.so: the B200 (sm_100a) implementation of the
 * Monte Carlo propagation hot path of otimgren/centrex-molecule-trajectories.
 *
 * The reference has no FFI of its own (it is pure Python); its plugin API for
 * this path is the abstract dataclass `BeamlineElement.propagate_through`
 * (src/trajectories/beamline_elements/apertures.py:22-54) driven by
 * `Beamline.propagate_through` (src/trajectories/beamline.py:20-38) inside the
 * per-molecule loop of `TrajectorySimulator.run_simulation`
 * (src/trajectories/trajectory_simulator.py:52-78).  Each entry point below
 * names the reference interface it replaces.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success or a negative CMT_E* code and never
 *     throws; cmt_last_error() returns a thread-local message for the last
 *     failure on the calling thread;
 *   - pointers documented "device" are caller-owned device memory on the
 *     handle's device (e.g. torch tensors' data_ptr()); the *_host entry
 *     points take host memory and do their own staging;
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed
 *     as void*, NULL = legacy default stream) and never synchronise;
 *   - a beamline handle is immutable after creation: concurrent launches on
 *     different streams are safe as long as each uses its own workspace and
 *     output buffers;
 *   - all arithmetic is IEEE-754 binary64 in the reference's operation order;
 *     fused multiply-adds and shared reciprocals appear only where the result
 *     is bit-identical to the separately rounded form (csrc/cmt_device.cuh,
 *     checked on the device by cmt_selftest).
 */
#ifndef CMT_H
#define CMT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMT_VERSION 101          /* 0.1.1: cmt_outputs_t.queue_capacity, cmt_resume */
#define CMT_MAX_ELEMENTS 40      /* element table lives in kernel-parameter constant memory */
#define CMT_MAX_FATES 64         /* fates are stored as uint8 and selected by a 64-bit mask */
#define CMT_MAX_TABLES 8
#define CMT_MAX_PLANES 16      /* probe planes per cmt_plane_crossings call */
#define CMT_WORK_SLOTS 8         /* length of the work-counter array */
#define CMT_ROW_DOUBLES 10       /* x,y,z,vx,vy,vz,ax,ay,az,t: one Trajectory row (molecule.py:133-144) */

/* error codes */
#define CMT_OK 0
#define CMT_EINVAL (-1)          /* bad argument */
#define CMT_ECUDA (-2)           /* CUDA runtime error (see cmt_last_error) */
#define CMT_ENOMEM (-3)          /* workspace too small / allocation failed */
#define CMT_ENODEV (-4)          /* no usable sm_100 device */

/* arithmetic modes (cmt_beamline_set_math) */
#define CMT_MATH_EXACT 0         /* default: every operation rounds as in the reference, results bit-identical */
#define CMT_MATH_CONTRACTED 1    /* same algorithm, fused multiply-adds and reciprocal multiplications:
                                  * agreement to ~1e-13 relative (1e-9 in the worst case, on coordinates that pass near zero), 1.3-1.6x the lens-integrator throughput */

/* element kinds */
#define CMT_CIRCULAR 0           /* CircularAperture,    apertures.py:83-115  */
#define CMT_RECTANGULAR 1        /* RectangularAperture, apertures.py:147-189 */
#define CMT_FIELDPLATES 2        /* FieldPlates,         apertures.py:213-270 */
#define CMT_LENS 3               /* ElectrostaticLens,   electrostatic_lens.py:23-118 */
#define CMT_HONEYCOMB 4          /* Honeycomb,           meshes.py:26-178 (cell centres as hexalattice.make_grid lays
                                  * them out, hit test as matplotlib's RegularPolygon.contains_point: both restated,
                                  * parity unpinned at that third-party boundary) */

/*
 * One flattened beamline element (replaces a BeamlineElement dataclass
 * instance; fields follow apertures.py:22-36,157-163,221-225 and
 * electrostatic_lens.py:29-46).  Elements must be sorted by z0
 * (beamline.py:40-45).
 */
typedef struct cmt_element {
    int32_t type;      /* CMT_CIRCULAR .. CMT_HONEYCOMB */
    int32_t fate;      /* fate id recorded on a hit; lens: id of "Lens entrance" */
    int32_t fate2;     /* lens only: id of "Inside lens" */
    int32_t table;     /* lens only: index into the tables given at creation */
    int32_t n_steps;   /* lens: int(rint(L/dz)), electrostatic_lens.py:87; honeycomb: nx (cells per row, meshes.py:50) */
    int32_t reserved;  /* honeycomb: ny (rows, meshes.py:51); 0 otherwise */
    double z0, z1;     /* entrance and exit planes, z1 = z0 + L */
    double x1, x2;     /* rectangular / field plates: open interval in x; honeycomb: x1 = make_grid's mid_x */
    double y1, y2;     /* rectangular: open interval in y; honeycomb: y1 = make_grid's mid_y */
    double R;          /* circular aperture / lens bore: d/2 (test is sqrt(x^2+y^2) > R about the origin);
                        * honeycomb: polygon radius (cell_wall_length*sqrt(3) - cell_wall_thickness/2)/2, meshes.py:73-77 */
    double dz;         /* lens: integration step along z; honeycomb: centre pitch cell_wall_length*sqrt(3), meshes.py:58 */
} cmt_element_t;

/* Lens radial-acceleration table a_r(r): what ElectrostaticLens.a_interp holds
 * (electrostatic_lens.py:32,209): r ascending, linear interpolation. Host memory. */
typedef struct cmt_table {
    const double *r;
    const double *a;
    int32_t n;
    int32_t reserved;
} cmt_table_t;

/* Molecular-beam source (replaces Distribution.draw, distributions.py:69-76,
 * 112-119,155-162).  Samples come from Philox4x32-10 keyed by `seed` and
 * indexed by the global molecule index (see DESIGN.md "Source"). */
#define CMT_POS_DISC 0           /* CeNTREXPositionDistribution: uniform disc of radius p0 */
#define CMT_POS_GAUSS 1          /* GaussianPositionDistribution: N(0,p0) x N(0,p1) */
typedef struct cmt_source {
    int32_t pos_kind;
    int32_t reserved;
    double vmean[3];
    double vsigma[3];
    double p0, p1;
    double z;
} cmt_source_t;

/* Output bundle of a propagation call; every pointer is device memory. */
typedef struct cmt_outputs {
    uint8_t *fate;          /* [n] fate id per molecule, or NULL */
    double *final_state;    /* [10][final_ld] SoA last trajectory row per molecule (x,y,z,vx,vy,vz,ax,ay,az,t), or NULL */
    int64_t final_ld;       /* leading dimension of final_state (>= n) */
    int64_t *counters;      /* [n_fates] per-fate counts, ACCUMULATED (Counter, trajectory_simulator.py:106-124); required */
    int64_t *work;          /* [CMT_WORK_SLOTS] accumulated: ballistic rows, lens RK steps, table out-of-range
                             * evaluations, lens entries, RK steps that took the plain-intrinsic path, molecules whose
                             * fate the FP32 filter of the walk kernel decided, molecules dropped by a full lens
                             * queue (see queue_capacity), 1 reserved; or NULL */
    int64_t *saved_index;   /* [saved_capacity] global indices of molecules whose fate is in save_mask (unordered), or NULL */
    int64_t *saved_count;   /* [1] accumulated cursor into saved_index (may exceed capacity: then the list is truncated) */
    int64_t saved_capacity;
    uint64_t save_mask;     /* bit f set: fate f is an "aperture of interest" (trajectory_simulator.py:75-76) */
    int64_t queue_capacity; /* 0 (default): the lens queue holds all n molecules of the launch.  k > 0: it holds k
                             * (workspace of cmt_workspace_bytes(bl, k) bytes); only ~0.5 % of a CeNTREX sample
                             * reaches the lens.  Molecules that find the queue full are DROPPED -- no fate, no
                             * Counter entry -- and counted in work[6]: a caller that sets this must pass `work`,
                             * read work[6] and repeat the launch with a larger capacity when it is not zero. */
} cmt_outputs_t;

typedef struct cmt_beamline cmt_beamline_t;

/* ---- lifetime ---------------------------------------------------------- */

/* Replaces building a `Beamline` (beamline.py:10-18): copies the element table
 * and lens tables to `device`, precomputes exact squared-radius thresholds and
 * table slopes.  fate ids must be < n_fates <= CMT_MAX_FATES. */
int cmt_beamline_create(const cmt_element_t *elements, int n_elements,
                        const cmt_table_t *tables, int n_tables,
                        int n_fates, int fate_detected, double g, int device,
                        cmt_beamline_t **out);
void cmt_beamline_destroy(cmt_beamline_t *bl);

/* Select the arithmetic mode of later launches on this handle (CMT_MATH_*).  Not to be called while
 * launches on the handle are being issued from another thread. */
int cmt_beamline_set_math(cmt_beamline_t *bl, int mode);

/* Rows a full trajectory can have: 1 + sum(N_steps) (molecule.py:115-131 without its 10 spare rows). */
int cmt_beamline_max_rows(const cmt_beamline_t *bl);
int cmt_beamline_device(const cmt_beamline_t *bl);

/* Bytes of device scratch a propagation call over up to n_max molecules needs: a 256-byte header of
 * counters and, for a beamline with a lens, two queue arrays of 64 B per molecule (worst case: every
 * molecule enters the lens); only the part that is used is ever touched. */
size_t cmt_workspace_bytes(const cmt_beamline_t *bl, int64_t n_max);

/* ---- the hot path ------------------------------------------------------ */

/* Replaces the per-molecule loop trajectory_simulator.py:62-76 for explicit
 * initial conditions: molecule i starts at (ic[0..2][i], ic[3..5][i]) with
 * a=(0,-g,0), t=0 (molecule.py:15-24) and is walked through every element
 * (beamline.py:20-38).  ic is device memory, SoA [6][ic_ld].  Global index of
 * molecule i (reported in saved_index) is first_index + i. */
int cmt_propagate_ic(const cmt_beamline_t *bl, int64_t n, int64_t first_index,
                     const double *ic, int64_t ic_ld, const cmt_outputs_t *out,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Same, with initial conditions generated on the device (replaces
 * vdist.draw/xdist.draw + the loop, trajectory_simulator.py:57-76). */
int cmt_propagate_philox(const cmt_beamline_t *bl, const cmt_source_t *src, uint64_t seed,
                         int64_t first_index, int64_t n, const cmt_outputs_t *out,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Materialise the source's samples: ic[c][j] for global index first_index + j
 * (index == NULL) or index[j] (device int64).  ic is device memory [6][ic_ld]. */
int cmt_philox_draw(const cmt_source_t *src, uint64_t seed, int64_t first_index,
                    const int64_t *index, int64_t n, double *ic, int64_t ic_ld, void *stream);

/* Full trajectories (replaces Trajectory.update bookkeeping, molecule.py:115-167)
 * for n selected molecules.  state is device memory SoA [n_comp][state_ld] with
 * n_comp = 6 (x,v; a and t default) or 10 (x,v,a,t: resume from an arbitrary
 * row, which is what BeamlineElement.propagate_through(molecule) needs).
 * select (device, optional): molecule j reads column select[j] - select_base.
 * rows:   device, or NULL to only count rows and report fates.  Without row_offset the
 *         layout is [n][max_rows][10] (rows past n_rows[j] are left untouched); with
 *         row_offset (device int64 [n], e.g. the exclusive scan of a counting call's n_rows)
 *         molecule j writes its rows compactly from row row_offset[j] on.
 * max_rows bounds the rows written per molecule; n_rows[j] always reports the full count. */
int cmt_trajectories(const cmt_beamline_t *bl, int64_t n, const double *state, int n_comp,
                     int64_t state_ld, const int64_t *select, int64_t select_base,
                     double *rows, int32_t max_rows, const int64_t *row_offset, int32_t *n_rows,
                     uint8_t *fate, void *stream);

/* Bulk form of BeamlineElement.propagate_through(molecule) for molecules that are already in flight
 * (apertures.py:38-42 called on a live Molecule; beamline.py:26-31 for the elements of this handle): molecule j
 * resumes from the row state[0..9][j] = x,y,z,vx,vy,vz,ax,ay,az,t (device, SoA [10][state_ld]; a_z must be 0) and
 * is walked through every element of the handle.  Outputs (device, each optional): last_row [10][last_ld] = the
 * last row of its trajectory (where it was stopped, or where the last element left it), n_rows[j] = rows it
 * would have recorded including the row it resumed from, fate[j] (the handle's "Detected" id = still alive).
 * This is the hand-back step for beamlines with elements implemented outside this library: the caller runs its
 * own element on the survivors of one handle and resumes them on the handle of the elements behind it. */
int cmt_resume(const cmt_beamline_t *bl, int64_t n, const double *state, int64_t state_ld,
               double *last_row, int64_t last_ld, int32_t *n_rows, uint8_t *fate, void *stream);

/* State of n selected molecules where they cross given z planes, without
 * storing trajectories: the device form of find_radial_pos_dist / find_vel_dist
 * (post_processing.py:20-140) — last row before the plane flown ballistically
 * with that row's stored acceleration, or the row itself when it lies on the
 * plane; a molecule whose last row is before the plane does not count (:43).
 * state / n_comp / state_ld / select / select_base as in cmt_trajectories.
 * z_planes: HOST array of n_planes <= CMT_MAX_PLANES values in ascending order.
 * out:   device [n_planes][5][out_ld] (x, y, vx, vy, vz of molecule j in column j);
 * valid: device [n_planes][out_ld], 1 where the molecule reached the plane
 *        (columns of out with valid == 0 are left untouched);
 * fate:  device [n] or NULL. */
int cmt_plane_crossings(const cmt_beamline_t *bl, int64_t n, const double *state, int n_comp,
                        int64_t state_ld, const int64_t *select, int64_t select_base,
                        const double *z_planes, int32_t n_planes, double *out, int64_t out_ld,
                        uint8_t *valid, uint8_t *fate, void *stream);

/* ---- host-buffer convenience (what a non-CUDA host language binds) ------ */

/* ic_host [6][n] (SoA, row-major), fate_host [n] or NULL, final_host [10][n]
 * or NULL, counters_host [n_fates] (accumulated), work_host [CMT_WORK_SLOTS] or NULL
 * (accumulated).  Stages through pinned buffers in chunks on two streams so
 * PCIe copies overlap the kernels; returns after everything has landed. */
int cmt_run_host_ic(const cmt_beamline_t *bl, int64_t n, const double *ic_host,
                    uint8_t *fate_host, double *final_host, int64_t *counters_host,
                    int64_t *work_host);

int cmt_run_host_philox(const cmt_beamline_t *bl, const cmt_source_t *src, uint64_t seed,
                        int64_t first_index, int64_t n, int64_t *counters_host,
                        int64_t *work_host);

/* ---- diagnostics ------------------------------------------------------- */

/* Per-kernel device time of the calls made on this thread since the last
 * reset, measured with CUDA events on the launch stream (enable first).
 * ms[0] = walk kernel (ballistic + aperture tests up to the first lens),
 * ms[1] = lens stage (the segment launches of the RK integrator + the tail kernel for the
 * elements behind the lens, timed as one interval per propagation call), ms[2] = trajectory
 * kernel, ms[3] = source-only kernel; launches[k] = number of timed intervals. */
int cmt_timing_enable(int on);
int cmt_timing_read(double ms[4], int64_t launches[4], int reset);
/* Measurement aid: with cmt_timing_enable(2) every kernel of the lens stage is bracketed by its own pair of events as
 * well; this call waits for all recorded intervals and returns them -- start and end in ms after the earliest start,
 * kind (0 walk kernel, 1 lens stage as a whole, 2 trajectory / resume kernels, 7 tail kernel, 8 + k lens segment k)
 * and a small id of the stream they ran on -- up to `capacity` entries; the return value is the number recorded.
 * The stage sums of cmt_timing_read are updated as well.  What overlapped with what, without a system profiler. */
int64_t cmt_timing_timeline(double *start_ms, double *end_ms, int32_t *kind, int32_t *stream_id, int64_t capacity);


/* Kernel launches this library has issued in the process so far (walk, lens segments, tail, source);
 * reset != 0 also sets the count back to zero. */
int64_t cmt_launch_count(int reset);

/* Measured FP64 pipe ceilings on `device` (dependent-chain-free DFMA / DADD
 * streams), in operations per second; used as roofline denominators. */
int cmt_fp64_peak(int device, double *dfma_per_s, double *dadd_per_s);

/* Arithmetic self-test on `device`: n pseudo-random operands compare the
 * shared-reciprocal division and the inline square root used by the kernels
 * with __ddiv_rn / __dsqrt_rn bit for bit.  mode 0: lens-integrator magnitudes,
 * 1: w / 6 in its two-operation form (exponents -396 .. 1023), 2: the whole binary64 range,
 * 3: a / sqrt(s) with the reciprocal taken from the square root's own iteration (the lens
 * force's two divisions) under the RK step's validity record.  out[0] quotients that took the
 * short sequence, out[1] mismatches among them, out[2]/out[3] the same for
 * square roots, out[4] mismatches of the division with fallback. */
int cmt_selftest(int device, int64_t n, uint64_t seed, int mode, int64_t out[5]);

/* Debug switch read by cmt_beamline_create: bit 0 forces the plain-intrinsic
 * arithmetic (no shared reciprocals), bit 1 switches the FP32 fate filter of the walk
 * kernel off, bit 2 keeps the filter but without its constant-threshold fast form, so
 * each variant can be compared with the plain one.
 * Returns the previous value; a negative argument only queries. */
int cmt_debug_flags(int flags);

int cmt_version(void);
const char *cmt_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CMT_H */
