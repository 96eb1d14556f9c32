// Native tick loop: GalaxySimulation.run (simulation.py:145-158) for the stock force, as ONE call.
// The reference dispatches ~23-55 ATen kernels per tick from Python; here a steady-state tick of a float mode is
// 2 launches — the kick-kick-drift kernel, which also reduces the j-split partial sums of the previous force pass
// and emits the packed sources, and the pair kernel — issued from C (int modes: reset / max-d² / level table /
// pair kernel / reduction with min-max, because the force snap needs the global extrema first).  For long runs
// of small systems the tick body is captured once into a CUDA graph and replayed, so the per-tick host cost is
// one cudaGraphLaunch.
#include "common.cuh"
#include "internal.cuh"
#include "stream.cuh"

using namespace nb;

namespace {

struct TickArgs {
    const void *x_in, *v_in; void* acc_in;     // state read by the FIRST tick (== x, v, acc for an in-place run)
    void *x, *v, *acc; const void* mass; int64_t n; int dim, dtype, mass_dtype, mode, levels, snap_levels;
    double G, eps_sq, min_dist_sq, dt; int uniform; double mass_value;
    void *packed, *table; int64_t* scalars; void* ws; int64_t ws_bytes;
};

// one tick body: [closing kick of the previous tick +] opening kick + drift (+ packed emit), then the force.
// Float modes (`deferred`): the pair kernel leaves partial sums in the workspace; their reduction is folded into
// the next tick's kick kernel (or done by finish_tick after the last one) — bit-identical to reducing first.
// `pe_out` != NULL (last tick of a call only): the force pass also accumulates the per-target potentials and the
// potential energy of the NEW positions is left in pe_out[0] (one more packed op per pair instead of a second O(N²) pass).
int enqueue_tick(const TickArgs& a, bool first, bool deferred, PartialSums* ps, cudaStream_t st, double* pe_out = nullptr) {
    const int phase = first ? NB_KDK_KICK_DRIFT : NB_KDK_KICK_KICK_DRIFT;
    int rc;
    if (deferred && !first)
        rc = kdk_from_partials(a.x, a.v, a.acc, a.x, a.v, a.n, a.dim, a.dtype, a.dt, phase, a.scalars, a.mass, a.mass_dtype, a.packed, 0,
                               *ps, st);
    else
        rc = nb_kdk(first ? a.x_in : a.x, first ? a.v_in : a.v, first ? a.acc_in : a.acc, a.x, a.v, a.n, a.dim, a.dtype, a.dt, phase,
                    first ? 0 : a.snap_levels, a.scalars, a.mass, a.mass_dtype, a.packed, 0, st);
    if (rc) return rc;
    if (a.levels > 0) {
        if ((rc = nb_reset_scalars(a.scalars, st))) return rc;
        // the force workspace is idle during pass 1: it doubles as the candidate buffer of the max-d² search
        if ((rc = nb_max_dist_sq(a.packed, a.n, a.dim, a.dtype, a.eps_sq, a.scalars, a.ws, a.ws_bytes, st))) return rc;
        if ((rc = nb_build_level_table(a.scalars, a.dtype, a.eps_sq, a.min_dist_sq, a.G, a.levels, a.table, st))) return rc;
    }
    PartialSums now{};
    if ((rc = accel_pairs(a.packed, a.n, a.x, a.n, a.dim, a.dtype, a.mode, a.G, a.eps_sq, a.table, a.levels, a.uniform, a.mass_value,
                          a.scalars, a.ws, a.ws_bytes, st, &now, pe_out != nullptr))) return rc;
    if (pe_out && (rc = potential_from_phi(now, a.dtype, a.mass, a.mass_dtype, a.eps_sq, pe_out, st))) return rc;
    if (deferred) {
        // the plan (split count, scale) is a pure function of the arguments: every tick produces the same descriptor,
        // which is what lets a captured tick body be replayed
        *ps = now;                                          // (packed_stale = false: the kick kernel above re-emitted the records)
        return NB_OK;
    }
    return accel_reduce(now, a.acc, a.scalars, st);
}

}  // namespace

extern "C" int nb_run_ticks(const void* x_in, const void* v_in, const void* acc_in, void* x, void* v, void* acc, const void* mass, int64_t n, int dim, int dtype, int mass_dtype, int mode,
                            int levels, int snap_levels, double G, double eps_sq, double min_dist_sq, double dt, int64_t ticks,
                            int uniform_mass, double mass_value, void* packed, void* level_table, int64_t* scalars,
                            void* workspace, int64_t workspace_bytes, int use_graph, double* pe_out, void* stream) {
    if (!x || !v || !acc || !mass || !packed || !scalars || !workspace || n <= 0 || ticks < 0) return NB_ERR_INVALID_ARGUMENT;
    if (ticks == 0) return NB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // the first kick only reads acc_in (it is already snapped, so no snap/write-back happens on it)
    TickArgs a{x_in ? x_in : x, v_in ? v_in : v, acc_in ? const_cast<void*>(acc_in) : acc, x, v, acc, mass, n, dim, dtype, mass_dtype, mode, levels, snap_levels, G, eps_sq, min_dist_sq, dt,
               uniform_mass, mass_value, packed, level_table, scalars, workspace, workspace_bytes};
    // the reduction can ride on the next kick kernel when no global min/max is needed between the two (no force
    // snap) and the accelerations have the state's dtype
    const bool acc_f64 = dtype == NB_F64 || mode == NB_MODE_FLOAT64;
    const bool deferred = snap_levels == 0 && acc_f64 == (dtype == NB_F64);
    if (pe_out && !((dtype == NB_F32 && mode == NB_MODE_FLOAT32) || (dtype == NB_F64 && mode == NB_MODE_FLOAT64)))
        return NB_ERR_UNSUPPORTED;                          // the fused potential needs the unquantised d² in the state dtype
    PartialSums ps{};
    int rc = enqueue_tick(a, /*first=*/true, deferred, &ps, st, ticks == 1 ? pe_out : nullptr);
    if (rc) return rc;
    int64_t remaining = ticks - 1;
    const int64_t tail = pe_out ? 1 : 0;                    // the last tick is enqueued directly when it carries the potential
    if (use_graph >= 2 && deferred && remaining - tail >= 2) {
        // small fp32 systems: all remaining ticks in ONE cooperative launch (grid barriers instead of launches)
        rc = persistent_ticks(x, v, acc, mass, mass_dtype, n, dim, dtype, mode, G, eps_sq, dt, uniform_mass, mass_value, packed, workspace,
                              workspace_bytes, remaining - tail, &ps, st);
        if (rc == NB_OK) remaining = tail;
        else if (rc != NB_ERR_UNSUPPORTED) return rc;       // unsupported / not co-resident: graph replay below
    }
    if (use_graph && remaining - tail >= 4) {
        // capture one steady-state tick on a private stream (the caller's may be the legacy default stream, which
        // cannot be captured), then replay it on the caller's stream
        cudaStream_t cap = nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = enqueue_tick(a, /*first=*/false, deferred, &ps, cap);
            e = cudaStreamEndCapture(cap, &graph);
            if (rc == NB_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        }
        if (rc == NB_OK && e == cudaSuccess) {
            for (; remaining > tail && e == cudaSuccess; --remaining) e = cudaGraphLaunch(exec, st);
        }
        if (exec) cudaGraphExecDestroy(exec);          // deferred until the in-flight launches finish
        if (graph) cudaGraphDestroy(graph);
        if (cap) cudaStreamDestroy(cap);
        if (rc) return rc;
        if (e != cudaSuccess) { cudaGetLastError(); return cuda_status(e); }
    }
    for (; remaining > 0; --remaining)
        if ((rc = enqueue_tick(a, /*first=*/false, deferred, &ps, st, remaining == 1 ? pe_out : nullptr))) return rc;
    // closing half kick (with the force snap of INT8/INT4) so that the state is observable; in the deferred case the
    // same kernel reduces the last force pass's partial sums and writes `acc`
    if (deferred && ps.packed_stale) {
        // the one-barrier kernel kept the packed records in shared memory: leave the records of the final positions behind,
        // as every other path does (the caller's potential-energy / force evaluations reuse them)
        if ((rc = nb_pack_sources(x, mass, n, dim, dtype, mass_dtype, packed, 0, st))) return rc;
    }
    if (deferred)
        return kdk_from_partials(nullptr, v, acc, nullptr, v, n, dim, dtype, dt, NB_KDK_KICK, scalars, mass, mass_dtype, nullptr, 0, ps, st);
    return nb_kdk(nullptr, v, acc, nullptr, v, n, dim, dtype, dt, NB_KDK_KICK, snap_levels, scalars, mass, mass_dtype, nullptr, 0, st);
}

extern "C" int nb_plan_splits(int64_t n_targets, int64_t n_chunks, int targets_per_block, int ctas_per_sm, int max_splits,
                              int* splits_out, int* chunks_per_split_out, int* target_blocks_out) {
    if (n_targets <= 0 || n_chunks <= 0 || targets_per_block <= 0 || max_splits <= 0) return NB_ERR_INVALID_ARGUMENT;
    const nb::SplitPlan p = nb::plan_splits(n_targets, n_chunks, targets_per_block, ctas_per_sm, max_splits);
    if (splits_out) *splits_out = p.splits;
    if (chunks_per_split_out) *chunks_per_split_out = p.chunks_per_split;
    if (target_blocks_out) *target_blocks_out = p.blocks_i;
    return NB_OK;
}
