// Library-internal entry points shared between translation units (hidden visibility; not part of the C ABI).
#pragma once
#include "common.cuh"

namespace nb {

// nb_accel split in two: the pair kernel, which leaves the j-split partial sums [splits][n_tgt·dim] (double) in
// `workspace`, and the reduction of those sums, acc[e] = (T)(Σ_s partial[s][e] · scale).  nb_run_ticks folds the
// reduction of tick t into the kick-kick-drift kernel of tick t+1 (one launch less per tick).
struct PartialSums {
    const double* partial;   // workspace
    int splits;
    int64_t count;           // n_tgt · dim
    double scale;            // G, G·m (uniform masses) or 1 (level-table factors carry G)
    bool out_f64;            // accelerations are fp64 (fp64 state or FLOAT64 mode)
    bool minmax;             // INT8/INT4: min/max of the outputs feed the force snap
    const double* partial_phi;   // [splits][n_tgt] per-target potentials Σ_j m_j / r_ij (PHI pass), else NULL
    double phi_scale;        // common mass (uniform-mass pass summed 1/r), else 1
    bool phi_uniform;
    int64_t n_tgt;
    bool packed_stale;       // the pass that produced these sums did not leave the packed records of its positions in global
                             // memory (one-barrier small-system kernel keeps them in shared memory): re-pack before reuse
};

// a contiguous window of the packed source set: chunks [first_chunk, first_chunk + n_chunks)
struct SourceWindow { int64_t first_chunk, n_chunks; int splits_before, max_splits; };

int accel_pairs(const void* packed_src, int64_t n_src, const void* pos_tgt, int64_t n_tgt, int dim, int dtype, int mode,
                double G, double eps_sq, const void* level_table, int levels, int uniform_mass, double mass_value,
                int64_t* scalars, void* workspace, int64_t workspace_bytes, cudaStream_t st, PartialSums* out,
                bool want_phi = false, const SourceWindow* window = nullptr);
// out[0] = Σ_{i<j} m_i m_j / r_ij (this shard's targets against all sources) from the potentials of a PHI pass
int potential_from_phi(const PartialSums& p, int dtype, const void* mass_tgt, int mass_dtype, double eps_sq, double* out,
                       cudaStream_t st);
int accel_reduce(const PartialSums& p, void* acc_out, int64_t* scalars, cudaStream_t st);

// `ticks` steady-state ticks of a small fp32 system in ONE cooperative launch (accel.cu); NB_ERR_UNSUPPORTED when out of scope
int persistent_ticks(void* x, void* v, void* acc, const void* mass, int mass_dtype, int64_t n, int dim, int dtype, int mode, double G,
                     double eps_sq, double dt, int uniform_mass, double mass_value, void* packed, void* workspace,
                     int64_t workspace_bytes, int64_t ticks, PartialSums* ps, cudaStream_t st);

// nb_kdk with the accelerations taken from j-split partial sums (reduced on the fly, written to `acc` as well)
int kdk_from_partials(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, int dim, int dtype,
                      double dt, int phase, const int64_t* scalars, const void* mass, int mass_dtype, void* packed_out,
                      int64_t total_chunks, const PartialSums& p, cudaStream_t st);

}  // namespace nb
