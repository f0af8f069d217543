/*
 * deepsolid_b200 -- C ABI of the B200-native local-energy hot path.
 *
 * The reference (bytedance/DeepSolid) has no FFI layer: its boundary is the set of
 * Python closures process.py builds (process.py:112-118,183-198).  Each entry point
 * below is what one of those closures computes for a whole batch of walkers, so
 * that deepsolid_b200/{network,hamiltonian,qmc,train}.py can re-create the closures
 * with the reference's names and argument meaning on top of this library.
 *
 * Conventions
 *  - every function returns 0 on success, a negative ds_status otherwise;
 *    ds_last_error() returns a thread-local human-readable message.
 *  - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 *  - "_dev" pointers are device memory on the context's device (owned by the
 *    caller, e.g. a torch tensor); entry points suffixed _host take HOST buffers
 *    and do the host<->device copies themselves on the given stream.
 *  - walkers are (B, 3N) fp64, electron-major then xyz, spin-up electrons first
 *    (network.py:322-323,451,537).
 *  - one context per (process, device); calls on one context are not thread safe
 *    (the reference driver is single-threaded, process.py:289-383).
 *  - there is no CPU fallback: without a CUDA device ds_ctx_create fails.
 */
#ifndef DEEPSOLID_B200_H
#define DEEPSOLID_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DS_API __attribute__((visibility("default")))
#else
#define DS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ds_ctx ds_ctx;

typedef enum ds_status {
    DS_OK = 0,
    DS_ERR_INVALID = -1,      /* invalid argument -> Python ValueError          */
    DS_ERR_CUDA = -2,         /* CUDA runtime error -> Python RuntimeError       */
    DS_ERR_UNSUPPORTED = -3,  /* option outside the implemented path            */
    DS_ERR_NOMEM = -4
} ds_status;

/* Laplacian evaluation modes of hamiltonian.local_energy_seperate (hamiltonian.py:209-218).
 * All four denote the same quantity; they select how the 3N directions are tiled. */
typedef enum ds_lap_mode {
    DS_LAP_FOR = 0, DS_LAP_HESSIAN = 1, DS_LAP_DIM_BATCH = 2, DS_LAP_PARTITION = 3
} ds_lap_mode;

/* Geometry + Ewald tables of one simulation cell.  All pointers are HOST memory and
 * are copied.  Replaces the attributes the hot path reads from the pyscf Cell
 * (supercell.py:64-140) and the tables EwaldSum.__init__ builds (ewaldsum.py:33-136). */
typedef struct ds_system_desc {
    int32_t n_up, n_dn;              /* simulation_cell.nelec                          */
    int32_t n_atoms_prim;            /* atoms of original_cell (network.py:643)        */
    int32_t n_atoms_sim;             /* atoms of the simulation cell (ewaldsum.py:41)  */
    const double *prim_latvec;       /* 3x3 row-major, rows = lattice vectors (Bohr)   */
    const double *sim_latvec;
    const double *prim_AV, *prim_BV; /* 3x3 (supercell.py:131-139, sym_type minimal)   */
    const double *sim_AV, *sim_BV;
    const double *prim_atoms;        /* n_atoms_prim x 3                               */
    const double *sim_atoms;         /* n_atoms_sim x 3                                */
    const double *sim_charges;       /* n_atoms_sim                                    */
    const double *klist_up;          /* n_up x 3 occupied k per orbital (hf.py:84-104) */
    const double *klist_dn;          /* n_dn x 3                                       */
    int32_t dist_kind;               /* 0 diagonal, 1 orthogonal, 2 general (distance.py:41-59) */
    const double *mi_shifts;         /* 27 x 3 minimal-image candidates (distance.py:64-67) */
    const double *lattice_displacements; /* 27 x 3 real-space images (ewaldsum.py:48-56) */
    double alpha;                    /* ewaldsum.py:63-64                              */
    int32_t n_g;
    const double *gpoints;           /* n_g x 3 (ewaldsum.py:68-89)                    */
    const double *gweight;           /* n_g                                            */
    const double *ion_exp_re, *ion_exp_im; /* n_g (ewaldsum.py:131-132)                */
    double ee_const, ei_const, ii_total;   /* ewaldsum.py:109-113,188-190              */
} ds_system_desc;

/* Network hyper-parameters (base_config.py:129-139).  Of the structural switches the
 * only use_last_layer=True is not implemented: every envelope_type, both distance functions, optional
 * orbital biases, per-spin or full determinants. */
typedef struct ds_net_desc {
    int32_t n_layers;      /* len(hidden_dims)            (3)   */
    int32_t hidden_one;    /* one-electron stream width   (256) */
    int32_t hidden_two;    /* two-electron stream width   (32)  */
    int32_t n_det;         /* determinants                (8)   */
    int32_t distance_type; /* 0 = 'nu' (network.py:189-224, 4 features per pair), 1 = 'tri' (network.py:227-246, 7) */
    int32_t envelope_type; /* 0 = isotropic, 1 = diagonal (sigma [A][3][n_s*D]), 2 = full (sigma [3][3][A][n_s*D]); network.py:335-364 */
    int32_t bias_orbitals; /* 1: every orbital[s] has a bias leaf b (2*n_s*D,) right after its w (network.py:177-179) */
    int32_t full_det;      /* 1: every spin channel makes N orbitals per determinant and ONE (N x N) determinant per k is
                            * taken (network.py:552-559): orbital w (H, 2*N*D), envelope (A, N*D), ds_orbitals -> (D, N, N) */
    int32_t use_last_layer;/* 1: `double` has n_layers entries and orbital w has 3*H + 2*P rows (network.py:129-134, 528-533).
                            * Implemented by every entry point; needs n_layers <= 3. */
} ds_net_desc;

DS_API const char *ds_last_error(void);
DS_API int ds_version(void);

DS_API int ds_ctx_create(const ds_system_desc *sys, const ds_net_desc *net, int device, ds_ctx **out);
DS_API int ds_ctx_destroy(ds_ctx *ctx);

/* Upper bound (bytes) for the internal per-chunk workspace; walkers are processed
 * in chunks that fit.  Default 16 GiB (env DS_WS_GIB overrides at context creation; DS_CHUNK_WALKERS caps the chunk directly). */
DS_API int ds_set_workspace_limit(ds_ctx *ctx, size_t bytes);

/* Parameter pytree of init_solid_fermi_net_params (network.py:135-184), flattened in
 * this order (L = n_layers):
 *   single[0].w, single[0].b, ..., single[L-1].w, single[L-1].b,
 *   double[0].w, double[0].b, ..., double[L-2].w, double[L-2].b,        (... double[L-1] with use_last_layer)
 *   orbital[0].w, [orbital[0].b,] orbital[1].w, [orbital[1].b,]      (b only with bias_orbitals)
 *   envelope[0].pi, envelope[0].sigma, envelope[1].pi, envelope[1].sigma
 * Every leaf is row-major fp64; pointers may be host or device memory; the data is
 * copied (and re-laid-out) on the legacy default stream, which is synchronised before the call returns, so the
 * caller may free or mutate it afterwards; device leaves produced on another stream must be complete before the call.
 *
 * Shapes this library implements (anything else is refused with DS_ERR_INVALID / DS_ERR_UNSUPPORTED at
 * ds_ctx_create; the reference accepts arbitrary hidden_dims, network.py:100-134):
 *   - 2 <= n_layers <= 4, the same (hidden_one, hidden_two) in every layer, hidden_two even and <= 32,
 *     use_last_layer only with n_layers <= 3, envelope isotropic / diagonal / full, at most 6 primitive-cell atoms with 'nu'
 *     features (layer-0 operand rows <= 32 columns);
 *   - the tcgen05 int8-slice path of the Laplacian sweep needs hidden_one and hidden_one + 2 hidden_two to be
 *     multiples of 64 and <= 512 (digit kernels: K % 8 == 0, K <= 512) and fewer than 2^31 Jacobian rows per chunk
 *     (the workspace limit keeps chunks far below that); other widths run the fp64 DMMA kernels;
 *   - the optional fused-digit sweep (DS_FUSED_DIGITS=1) additionally needs hidden_one == 256 and hidden_two <= 32. */
DS_API int ds_set_params(ds_ctx *ctx, const double *const *leaves, const int64_t *leaf_sizes, int n_leaves);

/* network.eval_func, methods eval_slogdet / eval_logdet / eval_phase_and_slogdet
 * (network.py:594-600), batched: log_abs[b] = log|psi|, phase[b] = angle(psi) in (-pi,pi].
 * Either output may be NULL. */
DS_API int ds_logpsi(ds_ctx *ctx, const double *x_dev, int64_t batch,
              double *log_abs_dev, double *phase_dev, void *stream);

/* Reverse-mode derivative of the batched network w.r.t. the parameters: what jax.jvp(batch_network, ...)
 * contributes to the energy-gradient estimator of train.make_loss.total_energy_jvp (train.py:129-137),
 *   tangents_dot = mean(Re(clip_diff * conj(d log psi)))  =  sum_b cot_abs[b] d log|psi_b| + cot_phase[b] d angle(psi_b)
 * with cot_abs = Re(clip_diff)/batch, cot_phase = Im(clip_diff)/batch.  `grad_leaves` holds n_leaves DEVICE
 * pointers in the leaf order and sizes of ds_set_params; every leaf is overwritten with its gradient. */
DS_API int ds_logpsi_vjp(ds_ctx *ctx, const double *x_dev, int64_t batch, const double *cot_abs_dev,
                  const double *cot_phase_dev, double *const *grad_leaves, const int64_t *leaf_sizes, int n_leaves,
                  void *stream);

/* Pullback through method eval_mats (network.py:601-602), what jax.grad of the pretraining loss
 * (pretrain.py:70-89) needs: cot_mats_dev has the layout of ds_orbitals' output,
 * grad_leaf = d/dleaf sum cot_re Re(M) + cot_im Im(M); leaves as in ds_logpsi_vjp. */
DS_API int ds_orbitals_vjp(ds_ctx *ctx, const double *x_dev, int64_t batch, const double *cot_mats_dev,
                    double *const *grad_leaves, const int64_t *leaf_sizes, int n_leaves, void *stream);

/* Plane-wave sums behind the observables of estimator.py (make_structure_factor :42-85: rho_q = sum_i exp(i q.x_i);
 * make_complex_polarization :15-40: exp(i sum_i G.x_i)).  q_dev [nq][3]; out_dev [batch][nq] complex (re, im);
 * mode 0 = sum of exponentials, 1 = exponential of the sum.  The batch means / pmean stay with the caller. */
DS_API int ds_rho_q(ds_ctx *ctx, const double *x_dev, int64_t batch, const double *q_dev, int nq, int mode,
             double *out_dev, void *stream);

/* Kronecker-factor statistics of the tagged dense layers: what the reference's KFAC estimator extracts from
 * total_energy_jvp (train.py:128-133; estimation mode fisher_exact, process.py:221; kfac_ferminet_alpha/estimator.py:
 * 284-320, tracer.py:196-332, curvature_blocks.py:262-281, DeepSolid/curvature_tags_and_blocks.py:142-156).  For every
 * register_repeated_dense layer (network.py:443), in the order single[0..L-1], double[0..L-2], orbital[spin 0, 1]:
 *   a_out[k] = sum_rows (x, 1)(x, 1)^T            [(in+1) x (in+1)], rows = walker x electron (x electron)
 *   g_out[k] = sum_rows ga ga^T + gp gp^T          [out x out]
 * ga / gp = cotangents of the layer output for d(sum_w log|psi_w|) / d(sum_w angle psi_w); the reference's output
 * factor is 2 g_out / rows (its tangent is sqrt(2)(ga - i gp)), its input factor a_out / rows (without the last row
 * and column when the layer has no bias).  env_abs / env_phase: n_env = 4 DEVICE leaves (pi_0, sigma_0, pi_1,
 * sigma_1) receiving the gradients of sum_w log|psi_w| and sum_w angle psi_w (the untagged envelope parameters get
 * a NaiveDiagonal block, curvature_blocks.py:111-133).  All outputs are raw sums over this call's batch. */
DS_API int ds_kfac_factors(ds_ctx *ctx, const double *x_dev, int64_t batch, double *const *a_out, const int64_t *a_sizes,
                    double *const *g_out, const int64_t *g_sizes, int n_layers, double *const *env_abs,
                    double *const *env_phase, const int64_t *env_sizes, int n_env, void *stream);

/* jax.value_and_grad(slog_network, argnums=1), batched (the `func` importance_update receives, qmc.py:325,101-118):
 * log|psi|, phase and d log|psi| / dx, d phase / dx of shape (batch, 3N).  Any output but one gradient may be NULL.
 * The gradients are the first-derivative half of the forward-Laplacian sweep (same cost as ds_local_energy). */
DS_API int ds_logpsi_grad_x(ds_ctx *ctx, const double *x_dev, int64_t batch, double *log_abs_dev, double *phase_dev,
                     double *grad_abs_dev, double *grad_phase_dev, void *stream);

/* method eval_mats (network.py:601-602): out = complex128 (re,im interleaved) of shape
 * (batch, 2 spins, n_det, n_s, n_s) with spin blocks concatenated (n_up block first). */
DS_API int ds_orbitals(ds_ctx *ctx, const double *x_dev, int64_t batch, double *out_dev, void *stream);
DS_API int64_t ds_orbitals_size(const ds_ctx *ctx); /* doubles per walker written by ds_orbitals */

/* hamiltonian.local_energy_seperate(f, cell, mode, partition_number)(params, x)
 * (hamiltonian.py:194-228), batched: kinetic = ke_re + i ke_im, ewald = ee+ei+ii. */
DS_API int ds_local_energy(ds_ctx *ctx, const double *x_dev, int64_t batch, int mode, int partition_number,
                    double *ke_re_dev, double *ke_im_dev, double *ewald_dev, void *stream);

/* EwaldSum.energy (ewaldsum.py:185-191): ee[b], ei[b] (constants included); ii via ds_ewald_ii. */
DS_API int ds_ewald(ds_ctx *ctx, const double *x_dev, int64_t batch, double *ee_dev, double *ei_dev, void *stream);
DS_API double ds_ewald_ii(const ds_ctx *ctx);

/* qmc.make_mcmc_step(...).mcmc_step (qmc.py:335-362) with mh_update (qmc.py:153-224),
 * symmetric all-electron Metropolis moves.
 *   x_dev        in/out walkers (batch, 3N)
 *   xi_dev,u_dev optional caller-supplied noise: gaussians (steps,batch,3N) and uniforms
 *                (steps,batch).  When NULL, a Philox4x32-10 stream keyed by `seed`
 *                generates them on the device.
 *   accept_dev   optional (steps,batch) uint8 accept masks
 *   n_accept_dev required: one double, number of accepted moves over all steps
 * pmove = n_accept / (steps*batch) is formed (and all-reduced) by the caller. */
DS_API int ds_mcmc_step(ds_ctx *ctx, double *x_dev, int64_t batch, int steps, double width, uint64_t seed,
                 const double *xi_dev, const double *u_dev, uint8_t *accept_dev,
                 double *n_accept_dev, void *stream);

/* train.make_loss.total_energy forward statistics (train.py:74-80): out6 =
 * [sum Re e_l, sum Im e_l, sum |e_l|^2, sum Re ke, sum ew, n] over the local batch;
 * the caller all-reduces the vector (NCCL) and forms mean / variance. */
DS_API int ds_energy_stats(ds_ctx *ctx, const double *ke_re_dev, const double *ke_im_dev,
                    const double *ewald_dev, int64_t batch, double *out6_dev, void *stream);

/* The hot path's collective: the statistics of `ds_energy_stats` and the Metropolis acceptance
 * reduced over the ranks with ONE ncclAllReduce(sum) of 8 doubles on `stream`.  Replaces
 * constants.pmean_if_pmap in train.py:78-80 (loss, imaginary, variance) and qmc.py:360-361 (pmove).
 *   comm            ncclComm_t as void* (NULL: single rank, identity).  NCCL is bound at run time from the
 *                   libnccl.so.2 the process has mapped; there is no link-time dependency.
 *   stats6_dev      output of ds_energy_stats on this rank
 *   n_accept_dev    optional: the accepted-move count of ds_mcmc_step; moves_per_rank = steps * batch_per_device
 *   global_variance 0 = the reference as written: mean over devices of (local <|e|^2> - |local <Re e>|^2)
 *                   (train.py:76-80 subtracts the LOCAL mean before the pmean); 1 = variance about the global mean
 *   out8_dev        [loss, imaginary, variance, mean Re ke, mean ewald, n_ranks, n_walkers, pmove]
 * ds_nccl_unique_id / ds_nccl_comm_init / ds_nccl_comm_destroy wrap ncclGetUniqueId / ncclCommInitRank /
 * ncclCommDestroy for hosts that have no NCCL binding of their own (rank 0 creates the 128-byte id and ships it to
 * the other ranks by whatever channel the host has). */
DS_API int ds_stats_allreduce(ds_ctx *ctx, void *comm, const double *stats6_dev, const double *n_accept_dev,
                       double moves_per_rank, int global_variance, double *out8_dev, void *stream);
DS_API int ds_nccl_unique_id(char *id128);
DS_API int ds_nccl_comm_init(void **comm, int n_ranks, const char *id128, int rank, int device);
DS_API int ds_nccl_comm_destroy(void *comm);

/* HOST-buffer forms: copy in, compute, copy out, synchronise the stream. */
DS_API int ds_logpsi_host(ds_ctx *ctx, const double *x_host, int64_t batch, double *log_abs_host, double *phase_host);
DS_API int ds_local_energy_host(ds_ctx *ctx, const double *x_host, int64_t batch, int mode, int partition_number,
                         double *ke_re_host, double *ke_im_host, double *ewald_host);
/* qmc.mh_one_electron_update (qmc.py:227-287) as driven by make_mcmc_step(one_electron_moves=True)
 * (qmc.py:355-358): steps * N single-electron moves, move i displaces electron i % N and re-wraps the walker.
 * xi_dev: optional gaussians (steps*N, batch, 3); u_dev, accept_dev: (steps*N, batch). */
DS_API int ds_mcmc_step_one_electron(ds_ctx *ctx, double *x_dev, int64_t batch, int steps, double width, uint64_t seed,
                              const double *xi_dev, const double *u_dev, uint8_t *accept_dev, double *n_accept_dev,
                              void *stream);

DS_API int ds_mcmc_step_host(ds_ctx *ctx, double *x_host, int64_t batch, int steps, double width, uint64_t seed,
                      const double *xi_host, const double *u_host, uint8_t *accept_host, double *n_accept_host);

/* ---- instrumentation (not part of the reference surface) ------------------------ */
/* number of kernels this library has launched on the context since creation */
DS_API int64_t ds_launch_count(const ds_ctx *ctx);
/* device-time (ms, CUDA events on the launching stream) and launch count of the
 * Jacobian-sweep GEMM kernel accumulated since the last reset; flops = executed DMMA flops */
DS_API int ds_profile_reset(ds_ctx *ctx);
DS_API int ds_profile_enable(ds_ctx *ctx, int on);
DS_API int ds_profile_get(ds_ctx *ctx, double *jac_ms, int64_t *jac_launches, double *jac_flops,
                   double *total_ms);
/* Copy an internal per-chunk buffer of the last ds_local_energy call (first chunk) to
 * dst_dev for stage-by-stage parity tests.  Returns number of doubles (or <0). */
/* walkers per chunk of the last batched call and the bytes of workspace held by the context */
DS_API int ds_workspace_info(ds_ctx *ctx, int64_t *chunk_walkers, int64_t *workspace_bytes);
DS_API int64_t ds_debug_buffer(ds_ctx *ctx, const char *name, double *dst_dev, int64_t max_doubles);
/* debug knobs: "stop_layer" = l stops ds_local_energy/ds_logpsi after one-electron layer l (-1: off);
 * "i8" = 0 runs the Jacobian-sweep GEMMs on the fp64 DMMA kernels instead of the tcgen05 int8-slice
 * kernels (default 1; environment DS_NO_I8=1 selects 0 at context creation). */
DS_API int ds_debug_set_int(ds_ctx *ctx, const char *key, int value);
/* Stand-alone run of the fp64 tensor-core GEMM kernel (C = A.B, row-major) for peak probes. */
DS_API int ds_dgemm_probe(int device, const double *a_dev, const double *b_dev, double *c_dev,
                   int64_t m, int n, int k, void *stream);

/* Stand-alone run of the tcgen05 (kind::i8, TMEM, TMA) sliced-integer fp64 GEMM (C = A.B, row-major).
 * gemm_ms: average device time of one GEMM launch over `reps`; slice_ms: digit kernel over A. */
DS_API int ds_ozaki_dgemm_probe(int device, const double *a_dev, const double *b_dev, double *c_dev,
                         int64_t m, int n, int k, int reps, double *gemm_ms, double *slice_ms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSOLID_B200_H */
