/*
 * nbody_b200.h — C ABI of libnbody_b200.so: the all-pairs softened-gravity leapfrog hot path of
 * nuclearbombmods/nbody-cosmological-simulation, as hand-written sm_100a CUDA.
 *
 * The reference has no FFI layer: its boundary is the Python surface of simulation.py /
 * quantization.py / metrics.py.  Each entry point below replaces the ATen op
 * stream of one reference function; the citation names the reference lines it stands in for.
 * The Python shell (nbody_cosmological_simulation_b200/) binds these with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions (binding for every function):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - no allocation, no host synchronisation, no global mutable state inside the library;
 *     scratch memory is passed in (`workspace`), sized by the nb_*_bytes() helpers (host-only);
 *   - return value: 0 = NB_OK, otherwise an NbStatus code; nb_error_string() explains it;
 *   - dtype codes: NB_F32 / NB_F64 — the dtype of the simulation state (positions/velocities);
 *   - `dim` is the spatial dimension D, 2 or 3; (n, D) arrays are row-major contiguous.
 */
#ifndef NBODY_B200_H_
#define NBODY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define NB_ABI_VERSION 2

typedef enum NbStatus {
    NB_OK = 0,
    NB_ERR_INVALID_ARGUMENT = 1,   /* bad dim / dtype / mode / size / null pointer / misalignment */
    NB_ERR_UNSUPPORTED = 2,        /* combination not implemented (said loudly, never a fallback) */
    NB_ERR_WORKSPACE_TOO_SMALL = 3,
    NB_ERR_NO_DEVICE = 4,          /* no sm_100 device / kernel image cannot run here */
    NB_ERR_CUDA_BASE = 1000        /* NB_ERR_CUDA_BASE + cudaError_t */
} NbStatus;

typedef enum NbDType { NB_F32 = 0, NB_F64 = 1 } NbDType;

/* quantization.py:10-18 PrecisionMode, same order */
typedef enum NbMode {
    NB_MODE_FLOAT64 = 0,
    NB_MODE_FLOAT32 = 1,
    NB_MODE_BFLOAT16 = 2,
    NB_MODE_FLOAT16 = 3,
    NB_MODE_INT8_SIM = 4,
    NB_MODE_INT4_SIM = 5,
    NB_MODE_CUSTOM = 6
} NbMode;

/* Slots of the device-side scalar block (int64[NB_SCALAR_SLOTS]).  Values are order-preserving
 * 64-bit keys of doubles (nb_key_from_double) so that cross-rank MIN/MAX all-reduces on the raw
 * int64 tensor are meaningful. */
enum {
    NB_SLOT_MAX_D2 = 0,     /* max over all pairs of d² (state dtype, widened)      quantization.py:113 */
    NB_SLOT_ACC_MIN = 1,    /* min over all n·D accelerations                      quantization.py:78  */
    NB_SLOT_ACC_MAX = 2,    /* max over all n·D accelerations                      quantization.py:79  */
    NB_SLOT_VAL_MIN = 3,    /* generic tensor min (nb_tensor_minmax)                                   */
    NB_SLOT_VAL_MAX = 4,    /* generic tensor max                                                     */
    NB_SLOT_RADIUS_MAX = 5, /* max radius                                           metrics.py:51      */
    NB_SCALAR_SLOTS = 8
};

/* ---- library / device ------------------------------------------------------------------- */
int nb_abi_version(void);
const char* nb_error_string(int status);
/* Fills SM count and compute capability of the current device; NB_ERR_NO_DEVICE if none. */
int nb_device_info(int* sm_count, int* cc_major, int* cc_minor);
int64_t nb_key_from_double(double v);
double nb_double_from_key(int64_t key);

/* ---- packed source set -------------------------------------------------------------------
 * The force / energy kernels stream sources from a chunk-major packed buffer that TMA bulk copies
 * (cp.async.bulk) move into shared memory.  One chunk = NB_CHUNK_UNITS units; a unit is two fp32
 * sources or one fp64 source (16 B of x,y + 16 B (D=3: z,m) or 8 B (D=2: m)).
 * Padding records fill the last chunk (and whole padding chunks when total_chunks asks for them): mass 0
 * and every coordinate = NB_PAD_COORD_F32 / NB_PAD_COORD_F64, far enough that d²^-3/2 underflows to exactly
 * 0, so a pad adds nothing to any force, potential or max-d² result. */
#define NB_PAD_COORD_F32 1.0e18f
#define NB_PAD_COORD_F64 1.0e150
#define NB_CHUNK_UNITS 128
int64_t nb_chunk_sources(int dtype);                       /* 256 for NB_F32, 128 for NB_F64 */
int64_t nb_chunk_bytes(int dim, int dtype);                /* 4096 (D=3) / 3072 (D=2)         */
int64_t nb_num_chunks(int64_t n, int dtype);               /* ceil(n / nb_chunk_sources)      */
int64_t nb_packed_bytes(int64_t n, int dim, int dtype);    /* nb_num_chunks * nb_chunk_bytes  */

/* Replaces the `pos.unsqueeze(0)` / `masses.unsqueeze(0)` operands of simulation.py:83,105,181-186.
 * pos: (n, dim) of `dtype`; mass: (n,) of `mass_dtype`; packed: nb_packed_bytes(n, dim, dtype).
 * total_chunks = 0 writes nb_num_chunks(n) chunks; a larger value also fills whole padding chunks
 * (equal-sized per-rank slices for the all-gather of an i-range-sharded run). */
int nb_pack_sources(const void* pos, const void* mass, int64_t n, int dim, int dtype, int mass_dtype,
                    void* packed, int64_t total_chunks, void* stream);

/* ---- force evaluation: GalaxySimulation._compute_accelerations, simulation.py:74-118 ------- */
/* Bytes of scratch nb_accel needs for n_targets targets (partial sums of the j-split). */
int64_t nb_accel_workspace_bytes(int64_t n_targets, int dim);
/* Split slots that workspace holds (<= 32 and <= 256 MiB of fp64 partial sums): the budget the windows of one windowed
 * evaluation (nb_accel_window) share. */
int nb_accel_max_splits(int64_t n_targets, int dim);
/* The j-split plan a pair-kernel launch would use (host arithmetic only, no launch): n_targets targets in blocks of
 * targets_per_block, n_chunks source chunks, ctas_per_sm resident CTAs per SM (148 SMs assumed when no device is present),
 * at most max_splits splits.  Test / tooling probe of the wave planner. */
int nb_plan_splits(int64_t n_targets, int64_t n_chunks, int targets_per_block, int ctas_per_sm, int max_splits,
                   int* splits_out, int* chunks_per_split_out, int* target_blocks_out);

/* INT8_SIM / INT4_SIM / CUSTOM pass 1 (quantization.py:112-113): max over ALL pairs of the n_src packed
 * sources of d² (the reference's exact rounding sequence, state dtype) -> scalars[NB_SLOT_MAX_D2] (atomic max;
 * reset it with nb_reset_scalars first).  Exact, but O(n) + O(C²): only sources in the outer shell of the point
 * set (R_i >= D_lb − R_max about the bounding-box centre) can form the farthest pair, and only those C candidates
 * are compared pairwise.  In a sharded run every rank holds the full source set and gets the global value; the
 * cross-rank MAX all-reduce of the slot is then a no-op kept for symmetry. */
int64_t nb_max_dist_workspace_bytes(int64_t n_src);
int nb_max_dist_sq(const void* packed_src, int64_t n_src, int dim, int dtype, double eps_sq, int64_t* scalars,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Bytes of the level table for `levels` grid levels (header + one 16-byte record per level + the
 * fast-lookup record). */
int64_t nb_level_table_bytes(int levels);
/* quantization.py:106-127 collapsed to a table: for each level k the snapped value u_k, the force
 * factor rn(rn(1/u_k^1.5)*G) (simulation.py:97-101) and the exact d² threshold at which
 * round((log(t)-lo)/(hi-lo)*(L-1)) steps from k-1 to k (bisection over the float bit pattern with
 * the reference's own op order).  lo = log(max(eps², min_dist_sq)) because the diagonal is part of
 * the tensor; hi comes from scalars[NB_SLOT_MAX_D2]. */
int nb_build_level_table(const int64_t* scalars, int dtype, double eps_sq, double min_dist_sq, double G,
                         int levels, void* table, void* stream);

/* Test hook for the fast level lookup nb_accel uses when levels <= 256 (csrc/lut.cuh): pushes EVERY float t
 * in [max(eps², min_dist_sq), max d²] through (a) quantization.py:106-120 evaluated op by op, (b) the fast lookup,
 * (c) its slow path, on the table nb_build_level_table wrote for the same scalars.  counters (uint64[4], device,
 * zeroed by the caller) += { floats tested, floats the fast lookup hands to the slow path ("doubt"),
 * fast-lookup levels != reference outside doubt, slow-path levels != reference }.  The last two must be 0. */
int nb_lut_selfcheck(const int64_t* scalars, double eps_sq, double min_dist_sq, int levels, const void* level_table,
                     uint64_t* counters, void* stream);

/* acc_out[i,:] = Σ_j f(q(d²_ij))·m_j·(x_j − x_i)  for targets pos_tgt[0:n_tgt] against all n_src
 * packed sources (self pairs contribute exactly 0, as `* (1 - eye)` at simulation.py:108).
 * mode selects q (quantization.py:21-71); int modes read `level_table`.
 * Output dtype: NB_F64 when dtype==NB_F64 or mode==NB_MODE_FLOAT64 (torch promotion at
 * quantization.py:45), else NB_F32.  For INT8/INT4 this is the PRE-snap acceleration; the min/max
 * over all outputs is folded into scalars[NB_SLOT_ACC_MIN/MAX] (quantization.py:78-79) and the snap
 * itself is nb_snap_accelerations or fused into nb_kdk.
 * uniform_mass != 0 asserts that every real source has mass == mass_value (the caller has checked): the
 * fp32-state FLOAT32 / FLOAT16 / BFLOAT16 kernels, the fp64-state FLOAT64 kernel and the level-table kernel for
 * levels <= 256 (INT8_SIM, INT4_SIM, CUSTOM) then drop the per-pair mass multiply (12 -> 11 packed fp32 ops,
 * 16 -> 15 fp64 ops per pair) and scale by mass_value once per target; other combinations ignore the hint. */
int nb_accel(const void* packed_src, int64_t n_src, const void* pos_tgt, int64_t n_tgt, int dim,
             int dtype, int mode, double G, double eps_sq, const void* level_table, int levels,
             int uniform_mass, double mass_value,
             void* acc_out, int64_t* scalars, void* workspace, int64_t workspace_bytes, void* stream);

/* nb_accel for FLOAT32 mode on fp32 state / FLOAT64 mode on fp64 state that ALSO returns the potential energy of the
 * same configuration: acc_out as nb_accel, pe_out[0] (device double) = ½ Σ_i m_i Σ_{j≠i} m_j / r_ij over the n_tgt
 * targets (an i-range shard adds its ranks; caller applies −G) — simulation.py:176-192 folded into the force pass
 * (get_total_energy() every tick, crash_point_test.py:190-197).  The targets must be among the sources (their j == i
 * term, d² == ε² exactly, is removed analytically).  mass_tgt: (n_tgt,) of mass_dtype. */
int nb_accel_potential(const void* packed_src, int64_t n_src, const void* pos_tgt, const void* mass_tgt, int64_t n_tgt,
                       int dim, int dtype, int mass_dtype, int mode, double G, double eps_sq, int uniform_mass,
                       double mass_value, void* acc_out, double* pe_out, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* Windowed force evaluation for i-range-sharded ticks (fp32 state in FLOAT32 mode, fp64 state in FLOAT64 mode).
 * nb_accel_window streams the contiguous source chunks [first_chunk, first_chunk + n_chunks) of the packed set with the
 * same kernels as nb_accel and APPENDS its j-split partial sums to `workspace` behind the `splits_before` split slots
 * earlier windows of the same evaluation wrote; *splits_total_out = splits_before + the slots it added (max_splits > 0 caps
 * them).  nb_accel_finish reduces all slots into acc_out exactly as nb_accel does.  A sharded tick can run the window of the
 * rank's OWN packed slot while the all-gather of the other ranks' slots is still in flight, then the windows over the
 * slots after and before its own (NB_B200_OVERLAP=1|2); measured on B200 the plain order — gather, then one nb_accel over
 * all slots — is faster and is what ShardedGalaxySimulation does by default. */
int nb_accel_window(const void* packed_src, int64_t n_src, int64_t first_chunk, int64_t n_chunks,
                    const void* pos_tgt, int64_t n_tgt, int dim, int dtype, int mode, double G, double eps_sq,
                    int uniform_mass, double mass_value, void* workspace, int64_t workspace_bytes, int splits_before,
                    int max_splits, int* splits_total_out, void* stream);
int nb_accel_finish(const void* workspace, int splits_total, int64_t n_tgt, int dim, int dtype, int mode, double G,
                    int uniform_mass, double mass_value, void* acc_out, void* stream);

/* Instrumentation: record the CUDA event `start_event` (cudaEvent_t as void*) immediately before the NEXT pair-kernel
 * launch issued by this host thread (inside nb_accel, nb_accel_potential, nb_accel_window or nb_run_ticks) and
 * `stop_event` immediately after the `launches`-th one from there (1 = the same launch; 2 = the two windows of a
 * sharded tick, which may sit on different streams), each on its launch's stream; one-shot.  bench.py times the dominant
 * kernel with it without leaving the default path. */
int nb_profile_next_force(void* start_event, void* stop_event, int launches);
/* Plain cudaEvent_t helpers for the hook above (so that a caller needs no CUDA runtime binding of its own):
 * create (timing enabled), elapsed milliseconds (synchronises on stop_event — host-blocking, instrumentation only),
 * destroy. */
int nb_event_create(void** event_out);
int nb_event_elapsed_ms(void* start_event, void* stop_event, float* ms_out);
int nb_event_destroy(void* event);

/* quantize_force -> _grid_quantize(a, levels) (simulation.py:115-116, quantization.py:74-88) with the
 * global min/max taken from scalars[NB_SLOT_ACC_MIN/MAX]; in place on acc (n*dim values, acc_dtype). */
int nb_snap_accelerations(void* acc, int64_t count, int acc_dtype, int levels, const int64_t* scalars,
                          void* stream);

/* ---- integrator: GalaxySimulation.step, simulation.py:120-143 ------------------------------ */
typedef enum NbKdkPhase {
    NB_KDK_KICK_DRIFT = 0,       /* v=v+a*(dt/2); x=x+v*dt                      simulation.py:132,135   */
    NB_KDK_KICK = 1,             /* v=v+a*(dt/2)                                simulation.py:141       */
    NB_KDK_KICK_KICK_DRIFT = 2   /* :141 of tick t fused with :132,135 of tick t+1 (same roundings)    */
} NbKdkPhase;
/* One HBM round trip of the state per tick.  mul and add are separately rounded (no FMA) and dt/2,
 * dt are cast to the state dtype first, exactly as torch does.  If snap_levels > 0 the acceleration
 * is first snapped to the linear grid (nb_snap_accelerations semantics) and written back to `acc`.
 * x_out/v_out may alias x_in/v_in.  If packed_out != NULL (phases with a drift) the packed source record of
 * every updated particle is emitted as well (mass: (n,) of mass_dtype; total_chunks as in nb_pack_sources). */
int nb_kdk(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, int dim,
           int dtype, double dt, int phase, int snap_levels, const int64_t* scalars,
           const void* mass, int mass_dtype, void* packed_out, int64_t total_chunks, void* stream);

/* GalaxySimulation.run for the stock force (simulation.py:145-158): `ticks` leapfrog ticks in one call.  The first
 * tick reads x_in, v_in, acc_in (the current state; acc_in already snapped, state dtype) and writes x, v, acc; later
 * ticks update x, v, acc in place.  Pass NULL for the *_in pointers to run fully in place.  The reference rebinds
 * its attributes to new tensors every tick: giving fresh x, v, acc buffers reproduces that without copies.  Per tick: nb_kdk(KICK_KICK_DRIFT, emitting `packed`) -> [int modes: nb_reset_scalars, nb_max_dist_sq,
 * nb_build_level_table] -> nb_accel; a closing nb_kdk(KICK) leaves a consistent, observable state.  Bit-identical
 * to issuing those calls one by one.  levels = d² grid levels (0 for float modes), snap_levels = force grid levels
 * (INT8/INT4, else 0).  use_graph = 1 captures the tick body once into a CUDA graph and replays it (worth it for
 * small systems where launch latency dominates); use_graph = 2 additionally runs the steady-state ticks of small fp32
 * systems in FLOAT32 mode (n <= 32768) as ONE persistent cooperative kernel with grid-wide barriers between the
 * integrator and force phases, falling back to the graph when the grid cannot be co-resident.  packed: nb_packed_bytes; level_table: nb_level_table_bytes (or
 * NULL); workspace: max(nb_accel_workspace_bytes, nb_max_dist_workspace_bytes) — the two uses never overlap.
 * pe_out (device double[1], may be NULL): when given, the force pass of the LAST tick also accumulates Σ_j m_j / r_ij
 * per target (one more packed op per source pair) and pe_out[0] receives Σ_{i<j} m_i m_j / r_ij of the final
 * positions — nb_potential_energy's value without its second O(N²) pass (simulation.py:176-192; caller applies −G).
 * Only where the pair loop sees the unquantised d² in the state dtype: fp32 state in FLOAT32 mode, fp64 state in
 * FLOAT64 mode; NB_ERR_UNSUPPORTED otherwise. */
int nb_run_ticks(const void* x_in, const void* v_in, const void* acc_in, void* x, void* v, void* acc,
                 const void* mass, int64_t n, int dim, int dtype, int mass_dtype,
                 int mode, int levels, int snap_levels, double G, double eps_sq, double min_dist_sq, double dt,
                 int64_t ticks, int uniform_mass, double mass_value, void* packed, void* level_table,
                 int64_t* scalars, void* workspace, int64_t workspace_bytes, int use_graph, double* pe_out,
                 void* stream);

/* ---- energies: simulation.py:170-196 ------------------------------------------------------- */
int64_t nb_energy_workspace_bytes(int64_t n_targets);
/* out[0] = this shard's part of Σ_{i<j} m_i m_j / sqrt(d²_ij) over UNORDERED pairs (double; the caller applies −G
 * and sums ranks).  The targets must be the sources tgt_offset .. tgt_offset+n_tgt−1 of the packed set (an i-range
 * shard).  With a chunk-aligned tgt_offset every unordered pair is evaluated once, by the half-ring rule: with C
 * chunks and δ = (source chunk − target chunk) mod C, a target takes a source chunk with weight 1 if 0 < 2δ < C and
 * ½ if δ == 0 or 2δ == C — half the pairs of the full matrix, and the same amount of work for every target chunk,
 * so equal-sized shards cost the same on every rank (the plain upper triangle gives rank 0 twice the mean).
 * A negative or unaligned offset selects the full-matrix form ½ Σ_{j≠i}, which needs no index correspondence. */
int nb_potential_energy(const void* packed_src, int64_t n_src, const void* pos_tgt, const void* mass_tgt,
                        int64_t n_tgt, int64_t tgt_offset, int dim, int dtype, int mass_dtype, double eps_sq,
                        double* out, void* workspace, int64_t workspace_bytes, void* stream);
/* out[0] = Σ_i m_i Σ_k v_ik²  (double; caller applies 0.5). */
int nb_kinetic_energy(const void* vel, const void* mass, int64_t n, int dim, int dtype, int mass_dtype,
                      double* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- rotation curve: metrics.py:25-78 ------------------------------------------------------ */
/* scalars[NB_SLOT_RADIUS_MAX] = max_i sqrt(Σ_k x_ik²) (metrics.py:48,51), atomic max. */
int nb_radius_max(const void* pos, int64_t n, int dim, int dtype, int64_t* scalars, void* stream);
/* Half-open bins [edges[b], edges[b+1]) in the state dtype; sum_vt[b] += |x·vy − y·vx| / max(r, 0.1),
 * count[b] += 1 (metrics.py:55-57,65-69).  sum_vt/count must be zeroed by the caller. */
int nb_rotation_curve(const void* pos, const void* vel, int64_t n, int dim, int dtype, const void* edges,
                      int num_bins, double* sum_vt, int64_t* count, void* stream);

/* ---- remainder of collect_metrics: metrics.py:81-95, 148-156 -------- */
int64_t nb_metrics_workspace_bytes(void);
/* compute_galaxy_radius: out[0] (state dtype, device) = k-th smallest (0-based) of sqrt(Σ_k x_ik²), found by
 * MSB-first radix select over the float bit patterns — bit-identical to torch.sort(radii)[0][k], no sort. */
int nb_radius_kth(const void* pos, int64_t n, int dim, int dtype, int64_t k, void* out, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* compute_velocity_dispersion: with s_i = |v_i| − |v_0| (shifted by the first star's speed for conditioning),
 * out[0] = Σ_i s_i, out[1] = Σ_i s_i² (doubles, device); the caller forms the unbiased standard deviation
 * sqrt((out[1] − out[0]²/n)/(n−1)), which is invariant under the shift. */
int nb_speed_moments(const void* vel, int64_t n, int dim, int dtype, double* out, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* compute_bound_fraction (metrics.py:98-145) without a sort and without gathering state.  The reference ranks the
 * stars by distance from the centre of mass, takes the cumulative mass in that order and calls a star bound when
 * |v| < sqrt(2 G M_enclosed / max(r, 0.1)).  Only the verdict is needed and it is monotone in M_enclosed, so a mass
 * histogram over nb_radius_bins() monotone radius bins brackets every star's enclosed mass; stars whose verdict is the
 * same at both ends of the bracket are counted at once, the few others ("doubt") get their exact enclosed mass from a
 * brute-force sweep.  Every stage is additive over i-range shards (all-reduce the arrays between the stages):
 *   1. nb_mass_moments           out[0..dim-1] = Σ m_i x_ik, out[dim] = Σ m_i                   metrics.py:118-119
 *   2. nb_radius_mass_histogram  hist[bin(r_i)] += m_i, r about `centre` (dim values, state dtype) metrics.py:122
 *   3. nb_exclusive_scan_f64     prefix[b] = Σ_{b' < b} hist[b']
 *   4. nb_bound_classify         counters[0] += stars bound for certain; counters[1] += stars in doubt, whose records
 *                                (nb_doubt_record_bytes() each) are appended to doubt_records (up to doubt_capacity;
 *                                size it for n to be safe); index_base = global index of local star 0 (tie order)
 *   5. nb_bound_resolve          partial[d] += Σ_{local j} m_j [r_j < r_d or (r_j == r_d and index_j <= index_d)]
 *   6. nb_bound_finish           counters[0] += doubt stars bound given their exact enclosed mass   metrics.py:133-139
 * counters: uint64[2], zeroed by the caller; hist / partial: doubles zeroed by the caller. */
int64_t nb_radius_bins(void);
int64_t nb_doubt_record_bytes(void);
int nb_mass_moments(const void* pos, const void* mass, int64_t n, int dim, int dtype, int mass_dtype, double* out,
                    void* workspace, int64_t workspace_bytes, void* stream);
int nb_radius_mass_histogram(const void* pos, const void* mass, const void* centre, int64_t n, int dim, int dtype,
                             int mass_dtype, double* hist, void* stream);
int nb_exclusive_scan_f64(const double* in, double* out, int64_t count, void* stream);
int nb_bound_classify(const void* pos, const void* vel, const void* mass, const void* centre, int64_t n,
                      int64_t index_base, int dim, int dtype, int mass_dtype, double G, const double* hist,
                      const double* prefix, uint64_t* counters, void* doubt_records, int64_t doubt_capacity,
                      void* stream);
int nb_bound_resolve(const void* pos, const void* mass, const void* centre, int64_t n, int64_t index_base, int dim,
                     int dtype, int mass_dtype, const void* doubt_records, int64_t n_doubt, double* partial,
                     void* stream);
int nb_bound_finish(const void* doubt_records, int64_t n_doubt, const double* enclosed, int dtype, int mass_dtype,
                    double G, uint64_t* counters, void* stream);
/* One 8-bit digit pass of nb_radius_kth's radix select, exposed for sharded runs: counts[d] += local radii whose key
 * matches `prefix` above bit shift+8 and has digit d at `shift` (a multiple of 8); all-reduce counts, pick the digit
 * that holds rank k, repeat towards shift 0.  counts: uint64[256], zeroed by the caller. */
int nb_radius_digit_histogram(const void* pos, int64_t n, int dim, int dtype, int shift, uint64_t prefix,
                              uint64_t* counts, void* stream);

/* ---- initial conditions at scale: galaxy.py:10-92, 142-211 as counter-based generators ----
 * Star i's random draws are Philox4x32-10(key = seed, counter = (i, stream)), so any partition of [0, num_stars) into
 * ranges [start, start+count) yields the same galaxy bit for bit: an i-range shard generates only its own slice.  Same
 * distributions and fp32 formulas as the reference; NOT torch's random stream (galaxy.create_disk_galaxy keeps that).
 * Global quantities are partition independent: the mean circular speed is accumulated as Σ round(v · nb_init_vsum_scale())
 * in int64 (exact under all-reduce), ranks in radius order come from a counting sort every rank can rebuild itself.
 *   nb_disk_galaxy_phase1      pos (count,2), noise-free tangential vel (count,2), mass (count,) — any may be NULL —
 *                              and vsum_fixed[0] += Σ round(v_circ · scale)                          galaxy.py:33-88
 *   nb_galaxy_add_dispersion   vel += N(0,1) · dispersion (Box-Muller on the star's Philox words; stream_id 0 for the
 *                              disk recipe, 1 for the halo recipe's second draw)                     galaxy.py:90,207
 *   nb_disk_radius_histogram   hist[bin(r_i)] += 1 over ALL stars (radii regenerated, nothing stored); nb_radius_bins() doubles
 *   nb_disk_radius_scatter     counting sort of all radii by bin: sorted_r / sorted_idx (num_stars each), cursor: uint32
 *                              per bin, zeroed; prefix = exclusive scan of hist
 *   nb_halo_phase1             local stars: exact rank inside their bin (ties by index) = enclosed visible mass,
 *                              + analytic NFW, circular speed, tangential vel, vsum_fixed            galaxy.py:176-204 */
double nb_init_vsum_scale(void);
int nb_disk_galaxy_phase1(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed,
                          int64_t start, int64_t count, float* pos, float* vel, float* mass, int64_t* vsum_fixed,
                          void* stream);
int nb_galaxy_add_dispersion(uint64_t seed, int stream_id, int64_t start, int64_t count, double dispersion, float* vel,
                             void* stream);
int nb_disk_radius_histogram(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed,
                             double* hist, void* stream);
int nb_disk_radius_scatter(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed,
                           const double* prefix, uint32_t* cursor, float* sorted_r, uint32_t* sorted_idx, void* stream);
int nb_halo_phase1(int64_t num_stars, double halo_radius, double dm_mass_ratio, int64_t start, int64_t count,
                   const float* pos, const double* hist, const double* prefix, const float* sorted_r,
                   const uint32_t* sorted_idx, float* vel, int64_t* vsum_fixed, void* stream);

/* ---- free-standing quantisers: quantization.py:74-127 -------------------------------------- */
int nb_reset_scalars(int64_t* scalars, void* stream);
/* scalars[VAL_MIN/VAL_MAX] = min/max of `in` (after clamp(min=clamp_min) and log() when log_space). */
int nb_tensor_minmax(const void* in, int64_t count, int dtype, int log_space, double clamp_min,
                     int64_t* scalars, void* stream);
/* _grid_quantize (quantization.py:74-88) with min/max from scalars[VAL_MIN/VAL_MAX]. */
int nb_grid_quantize(const void* in, void* out, int64_t count, int dtype, int levels, const int64_t* scalars,
                     void* stream);
/* _grid_quantize_safe (quantization.py:91-127) with log_min/log_max from scalars[VAL_MIN/VAL_MAX];
 * index_out (int32, may be NULL) receives the level index round(normalized). */
int nb_grid_quantize_safe(const void* in, void* out, int32_t* index_out, int64_t count, int dtype, int levels,
                          double min_val, const int64_t* scalars, void* stream);
/* The snap itself on identical pre-snap values: index_out[i] = round-half-even(normalized[i]). */
int nb_snap_index(const void* normalized, int32_t* index_out, int64_t count, int dtype, void* stream);
/* FLOAT16 / BFLOAT16 round trip of quantization.py:53,56 (mode = NB_MODE_FLOAT16 | NB_MODE_BFLOAT16). */
int nb_round_trip(const void* in, void* out, int64_t count, int dtype, int mode, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H_ */
