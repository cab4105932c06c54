/*
 * ccvm_b200.h -- C ABI of the B200-native CCVM dynamics engine (libccvm_b200.so).
 *
 * The reference (1QB-Information-Technologies/ccvm) is pure Python on torch; it has no FFI.
 * Every entry point below replaces ONE Python-level operator of the reference's hot path and is
 * what a ctypes binding inside the reference's solver classes would call (INTEGRATION.md shows
 * the stub).  Paths are relative to the reference's `ccvm_simulators/` package.
 *
 * Conventions
 *  - All matrices are fp32, row-major, resident in DEVICE memory unless the name says `_host`.
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).  All work is
 *    enqueued on it; nothing synchronises unless documented.
 *  - Return value: 0 on success, a negative CCVM_E_* code otherwise; `ccvm_last_error()` returns
 *    a thread-local message.  No CPU fallback exists: without a CUDA device every compute call
 *    fails with CCVM_E_CUDA.
 *  - y·Q means the ROW vector times the matrix, (yQ)_j = sum_i y_i Q_ij, as in the reference's
 *    einsum("bi,ij->bj") (solvers/dl_solver.py:145-149).
 */
#ifndef CCVM_B200_H
#define CCVM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCVM_ABI_VERSION 2

/* error codes */
#define CCVM_OK 0
#define CCVM_E_INVALID (-1)  /* bad argument / unsupported combination            */
#define CCVM_E_CUDA (-2)     /* CUDA runtime error (message has the CUDA string)   */
#define CCVM_E_TOO_LARGE (-3) /* problem does not fit the on-chip layout of any path */

/* solver ids: which reference class the loop belongs to */
#define CCVM_SOLVER_DL 0               /* solvers/dl_solver.py              DLSolver             */
#define CCVM_SOLVER_MF 1               /* solvers/mf_solver.py              MFSolver             */
#define CCVM_SOLVER_LANGEVIN 2         /* solvers/langevin_solver.py        LangevinSolver       */
#define CCVM_SOLVER_PUMPED_LANGEVIN 3  /* solvers/pumped_langevin_solver.py PumpedLangevinSolver */

/* algorithm ids */
#define CCVM_ALG_ORIGINAL 0 /* Solver._solve      */
#define CCVM_ALG_ADAM 1     /* Solver._solve_adam */

/* noise source */
#define CCVM_RNG_PHILOX 0 /* in-kernel generation keyed by (seed, offset, GLOBAL trajectory, variable): Philox4x32-10 per
                             (trajectory, step, 4 variables) on the tcgen05 path; on the SIMT paths Philox4x32-10 seeds
                             one xoshiro128+ stream per (trajectory pair, 4 variables); Box-Muller on both.
                             ccvm_dump_noise writes exactly the normals this mode draws */
#define CCVM_RNG_REPLAY 1 /* validation: consume a recorded noise tensor [T][K][N][B]                 */

/* post-processor ids (post_processor/factory.py:12-35; only the two batched ones are on the hot path) */
#define CCVM_PP_NONE 0
#define CCVM_PP_GRAD_DESCENT 1 /* post_processor/grad_descent.py:58-64 */
#define CCVM_PP_ADAM 2         /* post_processor/adam.py:58-66 (effective 1-step behaviour) */

/*
 * One call of Solver._solve / Solver._solve_adam.
 *
 * Replaces: DLSolver._solve (dl_solver.py:468-569), DLSolver._solve_adam (571-769),
 *           MFSolver._solve (mf_solver.py:493-593), MFSolver._solve_adam (595-764),
 *           LangevinSolver._solve (langevin_solver.py:368-435), ._solve_adam (437-561),
 *           PumpedLangevinSolver._solve (pumped_langevin_solver.py:232-309), ._solve_adam (311-449),
 *           including the hooks they call (calculate_drift / calculate_grads / fit_to_constraints).
 *
 * The whole iteration loop runs as ONE persistent kernel launch.
 */
typedef struct ccvm_solve_desc {
  int32_t solver;     /* CCVM_SOLVER_*                                                        */
  int32_t algorithm;  /* CCVM_ALG_*                                                           */
  int32_t n;          /* problem_size                                                         */
  int32_t batch;      /* batch_size: trajectories solved by THIS call                         */
  int32_t iterations; /* parameter_key[n]["iterations"]                                       */
  int32_t pump_rate_flag; /* __call__(pump_rate_flag=...) ; ignored by Langevin               */
  const float* q;     /* device [n*n]  self.q_matrix (already scaled & negated by the loader) */
  const float* v;     /* device [n]    self.v_vector                                          */
  double lower;       /* instance.solution_bounds[0]                                          */
  double upper;       /* instance.solution_bounds[1]                                          */
  /* Saturation value S exactly as the reference hands it to _solve/_solve_adam: DL = the
   * constructor's S, others = parameter_key[n]["S"].  If `s_vec` is non-NULL it is a device
   * [n] per-variable S (the reference's 1-D tensor S broadcast over the batch,
   * dl_solver.py:843-848) and `s` is ignored wherever the reference would use the tensor. */
  double s;
  const float* s_vec;
  /* parameter_key / __call__ scalars; unused ones are ignored per solver */
  double pump, dt, noise_ratio, j, sigma, feedback_scale, g;
  /* AdamParameters.to_dict() (solvers/algorithms.py:38-45); used when algorithm == ADAM */
  double alpha, beta1, beta2;
  int32_t add_assign;
  /* noise */
  int32_t rng_mode;    /* CCVM_RNG_*                                                          */
  const float* noise;  /* REPLAY: device [iterations][K][n][noise_batch], K = 2 for DL else 1  */
  int64_t noise_batch; /* REPLAY: trajectory extent of `noise` (>= traj_base + batch)          */
  uint64_t seed;       /* PHILOX key                                                          */
  uint64_t offset;     /* PHILOX stream offset (advance by >= 1 between calls)                */
  int64_t traj_base;   /* global index of this call's trajectory 0 (batch sharding across GPUs:
                          noise depends on the GLOBAL trajectory index, so results do not
                          depend on how the batch was split).  Must be EVEN in PHILOX mode:
                          noise streams belong to trajectory pairs (2p, 2p+1)               */
  /* outputs, device [batch*n] each, row-major (B,N):
   *   DL:  out0 = c (clamped), out1 = s
   *   MF:  out0 = mu, out1 = mu_tilde (clamped last measurement), out2 = sigma
   *   Langevin / PumpedLangevin: out0 = c                                                   */
  float* out0;
  float* out1;
  float* out2;
  /* optional evolution sampling (dl_solver.py:557-564): when evolution_step > 0, state is
   * stored whenever i % evolution_step == 0 or i + 1 >= iterations into
   * samples[k][sample][b][n], k over (c,s) / (mu,sigma) / (c). */
  int32_t evolution_step;
  int32_t num_samples;
  float* samples;
} ccvm_solve_desc;

int ccvm_solve(const ccvm_solve_desc* desc, void* stream);

/*
 * Validation aid, no reference counterpart: writes the standard normals that ccvm_solve(desc) draws in
 * CCVM_RNG_PHILOX mode (same seed / offset / traj_base / n / batch / iterations / solver) into
 * noise[iterations][K][n][batch] (device, K = 2 for DL else 1) -- the layout CCVM_RNG_REPLAY consumes and
 * the layout in which the reference's per-iteration draws are recorded (dl_solver.py:512-519:
 * Normal(0,1).sample((N,)).T, one [n][batch] slab per draw).  Replaying this tensor through the
 * reference arithmetic reproduces a production run, which pins the production kernel variants
 * (in-loop noise generation, compile-time column-group counts) to the oracle.
 * Which generator a launch uses depends on the kernel family the library picks for it (noise streams per
 * (trajectory pair, 4 variables) in the tiled kernels, per (trajectory pair, variable) in the small-n tensor-core
 * kernel, counter mode on the n > 256 tensor-core path), and that choice depends on the batch of the LAUNCH: to
 * dump a slice of a larger launch (batch trajectories from traj_base), put the launch's batch into
 * desc->noise_batch (0: the launch is desc->batch trajectories).
 */
int ccvm_dump_noise(const ccvm_solve_desc* desc, float* noise, void* stream);

/*
 * Many instances in ONE launch (grid over instances x trajectory blocks).
 *
 * Replaces: the user-level loop over instance files around Solver.__call__
 *           (examples/ccvm_boxqp_dl.py:27-52 and siblings).  `descs[i]` are
 * ordinary ccvm_solve descriptors that may differ in n, batch, iterations, scalars and pointers but
 * must share solver and algorithm, use CCVM_RNG_PHILOX and no evolution sampling.  Instances with
 * n <= 128 are bucketed by CTA size and solved by one launch per bucket; larger ones fall back to
 * one ccvm_solve each.  Results are identical to calling ccvm_solve on every descriptor.
 */
int ccvm_solve_batch(const ccvm_solve_desc* descs, int32_t count, void* stream);

/*
 * The tail of Solver.__call__: optional affine change of variables, optional batched
 * post-processor, optional second change of variables (the DL double map), BoxQP energy.
 *
 * Replaces: CCVMSolver.change_variables (dl_solver.py:219-235 and siblings; Langevin's
 *           (c+S)/(2S), langevin_solver.py:717-722), PostProcessorGradDescent.postprocess
 *           (post_processor/grad_descent.py:13-68), PostProcessorAdam.postprocess
 *           (post_processor/adam.py:15-69), ProblemInstance.compute_energy
 *           (problem_classes/boxqp/problem_instance.py:226-241).
 *
 *   x  = state * map1_scale_j + map1_shift          (map1_scale_vec overrides the scalar)
 *   pv = post_process(x)                            (or x)
 *   cf = pv * map2_scale_j + map2_shift             (skipped when apply_map2 == 0: cf = pv)
 *   energy_b = (0.5 * cf_b Q cf_b + V . cf_b) * scaled_by
 */
typedef struct ccvm_epilogue_desc {
  int32_t n;
  int32_t batch;
  const float* q;
  const float* v;
  const float* state; /* device [batch*n]                                              */
  int32_t apply_map1;
  double map1_scale, map1_shift;
  const float* map1_scale_vec; /* optional device [n]                                  */
  int32_t post_processor;      /* CCVM_PP_*                                            */
  int32_t pp_iterations;       /* grad-descent: num_iter_pp (reference default 10)     */
  double pp_step;              /* grad-descent step_size (0.1); adam lr (0.01)         */
  double pp_lower, pp_upper;   /* clamp bounds (0, 1)                                  */
  int32_t apply_map2;
  double map2_scale, map2_shift;
  const float* map2_scale_vec;
  double scaled_by;            /* instance.scaled_by                                   */
  float* problem_variables;    /* out, device [batch*n] (pv); may be NULL              */
  float* energy;               /* out, device [batch]; may be NULL                     */
  const float* scaled_by_dev;  /* optional DEVICE scalar that overrides scaled_by: the factor written by
                                  ccvm_scaling_factor can stay on the device (no host read-back, so a sweep
                                  never synchronises while it plans the next chunk); NULL: use scaled_by */
} ccvm_epilogue_desc;

int ccvm_epilogue(const ccvm_epilogue_desc* desc, void* stream);

/*
 * One whole Solver.__call__ as ONE kernel launch: the per-iteration schedules, the iteration loop,
 * the change of variables, the post-processor, the BoxQP energy and the solution statistics
 * (dl_solver.py:771-999 and siblings, including solution.py:65-146).  Every CTA of the persistent kernel
 * finishes its own trajectories (same device code as ccvm_epilogue, so the values are identical) and
 * merges best / argmin / the seven success counters across CTAs with a handful of atomics.
 *
 * `solve` is an ordinary ccvm_solve descriptor (its outputs are still written); `epi` supplies the maps,
 * the post-processor and the output buffers problem_variables (optional) and energy (required when
 * `result` is given) -- its q, v, state, n and batch are ignored (taken from `solve`: the state is out0,
 * or out1 = mu_tilde for MF).  `result` (device, 56 bytes, may be NULL) receives
 *   { float best; int32 arg_best; int32 counts[7]; uint32 ctas; uint64 loop_ns; uint64 tail_ns }
 * i.e. the block of ccvm_solution_stats followed by the number of CTAs and the device-measured duration
 * (max over CTAs) of the loop and of the tail.  Paths that cannot fuse (the tcgen05 kernels) run the
 * stand-alone kernels behind the same call; loop_ns / tail_ns are 0 then.
 */
int ccvm_solve_fused(const ccvm_solve_desc* solve, const ccvm_epilogue_desc* epi, double optimal_value,
                     void* result, void* stream);

/* ccvm_solve_batch with the fused tail of ccvm_solve_fused: `epis[i]` / `optimal_values[i]` (host arrays)
 * belong to `descs[i]`, `results` is device memory of count x 56 bytes (may be NULL).  A sweep chunk is
 * one launch per bucket instead of 1 + count epilogues + 1 statistics kernel. */
int ccvm_solve_batch_fused(const ccvm_solve_desc* descs, const ccvm_epilogue_desc* epis,
                           const double* optimal_values, int32_t count, void* results, void* stream);

/* Replaces ProblemInstance.compute_energy (problem_instance.py:226-241). */
int ccvm_compute_energy(const float* x, const float* q, const float* v, double scaled_by,
                        int32_t batch, int32_t n, float* energy, void* stream);

/* Replaces PostProcessorGradDescent.postprocess (grad_descent.py:58-64); x is updated in place. */
int ccvm_postprocess_grad_descent(float* x, const float* q, const float* v, int32_t batch,
                                  int32_t n, int32_t iterations, double step_size, double lower,
                                  double upper, void* stream);

/* Replaces PostProcessorAdam.postprocess (adam.py:58-66, one Adam step + clamp); in place. */
int ccvm_postprocess_adam(float* x, const float* q, const float* v, int32_t batch, int32_t n,
                          double lr, double lower, double upper, void* stream);

/*
 * Replaces Solution.__post_init__ / get_solution_stats (solution.py:65-146):
 *   best = max_b(-energy_b), arg_best = its index, counts[k] = #{b : gap_b <= thr_k},
 *   gap_b = (optimal - (-energy_b)) * 100 / |energy_b|, thr = {0.1, 1, 2, 3, 4, 5, 10}.
 * `result` is a device buffer of 9 x 4 bytes: {float best, int32 arg_best, int32 counts[7]}.
 */
int ccvm_solution_stats(const float* energy, int32_t batch, double optimal_value, void* result,
                        void* stream);

/*
 * ccvm_solution_stats for `count` instances in one launch (one 36-byte record each): the per-result
 * Solution construction of a sweep (examples/ccvm_boxqp_plot.py:40-82 builds one Solution per
 * instance, each with eight .item() syncs) becomes one kernel and one device->host copy.
 * `energy` is the concatenation of the instances' energy vectors, instance i occupying
 * [offsets[i], offsets[i+1]) (device int64[count+1]); `optimal_values` is device float[count];
 * `result` is device memory of count x 9 x 4 bytes laid out as in ccvm_solution_stats.
 */
int ccvm_solution_stats_batch(const float* energy, const int64_t* offsets,
                              const float* optimal_values, int32_t count, void* result,
                              void* stream);

/*
 * Multi-GPU result records (no reference counterpart: the reference is single-device).
 * A record is CCVM_RECORD_HEADER + n 32-bit words in a float buffer:
 *   [0] min energy (f32)   [1],[2] global index of the winner, low / high 32 bits (integers)
 *   [3..9] the 7 success counters (int32)   [10..] the winner's solution vector (f32 x n).
 * ccvm_pack_record: from the 9-word block written by ccvm_solution_stats (or the head of a
 *   ccvm_solve_fused result) and the local solution matrix problem_variables[batch][n].
 * ccvm_merge_records: reduces world_size such records (gathered[world_size][HEADER + n], e.g. the
 *   output of an all-gather) to one whose slot [0] is the best objective = max(-E); counters are
 *   summed, ties go to the lowest rank, NaN propagates.
 */
#define CCVM_RECORD_HEADER 10
int ccvm_pack_record(const void* stats, const float* problem_variables, int32_t n, int64_t traj_base,
                     float* record, void* stream);
int ccvm_merge_records(const float* gathered, int32_t world_size, int32_t n, float* merged,
                       void* stream);

/* Replaces CCVMSolver.get_scaling_factor (solvers/ccvm_solver.py:134-150):
 *   *out (device float) = sqrt(sum |Q_ij|) * multiplier. */
int ccvm_scaling_factor(const float* q, int32_t n, double multiplier, float* out, void* stream);

/*
 * The reference's per-solver operator hooks evaluated on arbitrary inputs (the "de-facto
 * operator API" its tests exercise, tests/unit/solvers/test_mf_solver.py:63-204).
 *
 * Replaces: _calculate_drift_boxqp / _calculate_grads_boxqp of DLSolver (dl_solver.py:117-217),
 *           MFSolver (mf_solver.py:141-233), LangevinSolver (langevin_solver.py:117-166),
 *           PumpedLangevinSolver (pumped_langevin_solver.py:95-147).
 * kind 0 = calculate_grads, 1 = calculate_drift.  Inputs/outputs are device (batch, n) tensors:
 *   DL   in0 = c, in1 = s                -> out0, out1 (c and s parts)
 *   MF   drift: in0 = mu, in1 = mu_tilde, in2 = sigma -> out0 = drift_mu, out1 = drift_sigma
 *        grads: in0 = mu_tilde                       -> out0
 *   Langevin / PumpedLangevin: in0 = c   -> out0
 * `pump` is the hook's pump argument (MF: instantaneous pump; PumpedLangevin: p), `rate` the DL
 * pump-rate multiplier.
 */
typedef struct ccvm_hook_desc {
  int32_t solver;
  int32_t kind;
  int32_t n;
  int32_t batch;
  const float* q;
  const float* v;
  const float* in0;
  const float* in1;
  const float* in2;
  double lower, upper;
  double s;
  const float* s_vec; /* optional device [n] */
  double pump, rate, feedback_scale, j, g;
  float* out0;
  float* out1;
} ccvm_hook_desc;

int ccvm_eval_hook(const ccvm_hook_desc* desc, void* stream);

/* Replaces CCVMSolver.change_variables (dl_solver.py:219-235): out = 0.5*x/S*(u-l) + 0.5*(u+l),
 * evaluated in the reference's operation order.  s_vec optional device [n]. */
int ccvm_change_variables(const float* x, float* out, int32_t batch, int32_t n, double lower,
                          double upper, double s, const float* s_vec, void* stream);

/* Replaces CCVMSolver.fit_to_constraints (dl_solver.py:237-250): out = clamp(x, lo, hi).
 * Bounds are scalars, or device tensors of `bound_len` = n (per variable) or batch*n elements. */
int ccvm_fit_to_constraints(const float* x, float* out, int32_t batch, int32_t n, double lo,
                            double hi, const float* lo_t, const float* hi_t, int64_t bound_len,
                            void* stream);

/* Replaces ProblemInstance.scale_coefs (problem_instance.py:243-255): q_out = q / f, v_out = v / f.
 * `factor` is a device tensor of factor_len = 1 (scalar) or n*n elements; in the latter case v is
 * broadcast against it like torch does and v_out has n*n elements. */
int ccvm_scale_coefs(const float* q, const float* v, int32_t n, const float* factor,
                     int64_t factor_len, float* q_out, float* v_out, void* stream);

/* Replaces the instance source of a synthetic sweep (no reference function: its instances are files,
 * problem_classes/boxqp/problem_instance.py:116-224): fills q[n*n] and v[n]
 * (device) with a dense symmetric BoxQP instance whose coefficient statistics match the bundled
 * benchmarking instances (off-diagonal std q_offdiag_std, diagonal std sqrt(2) x that, V std v_std),
 * in the reference's in-memory sign convention.  Deterministic in (n, seed). */
int ccvm_generate_boxqp(float* q, float* v, int32_t n, uint64_t seed, double q_offdiag_std,
                        double v_std, void* stream);

/*
 * Host-buffer convenience used for end-to-end timing: copies Q and V from HOST memory,
 * runs ccvm_solve_fused (one launch) on `stream`, copies energy[batch] and the
 * 9-word stats block back to HOST memory and synchronises the stream.  `solve` and `epi` carry
 * the scalar parameters; their device pointers (q, v, state, outputs) are ignored and replaced
 * by internal buffers.  `h_*` pointers are host memory (pinned recommended).
 */
int ccvm_solve_host(const ccvm_solve_desc* solve, const ccvm_epilogue_desc* epi,
                    const float* h_q, const float* h_v, double optimal_value, float* h_energy,
                    void* h_stats, void* stream);

/* Register-only FP32 FMA throughput probe (roofline denominator for the SIMT path).
 * mode 0 = scalar FFMA, mode 1 = packed FFMA2 (fma.rn.f32x2).  Runs on `stream`, synchronises,
 * writes achieved TFLOP/s. */
int ccvm_microbench_fp32(int32_t mode, double* tflops, void* stream);

/* Roofline denominator of the tensor-core path, no reference counterpart: measured rate of
 * back-to-back tcgen05.mma kind::tf32 (M = 128 per CTA, N = 256, K = 8, operands resident in
 * shared memory, FP32 accumulators in TMEM) over all SMs, in dense TF32 TFLOP/s.
 * mode 1: cta_group::1, mode 2: cta_group::2 (CTA pairs).  Synchronises the stream. */
int ccvm_microbench_tf32(int32_t mode, double* tflops, void* stream);

/* Query the launch geometry ccvm_solve would use: fills threads per CTA, CTAs, trajectories per
 * CTA, dynamic shared memory bytes, registers per thread.  For reports and tests. */
int ccvm_query_launch(const ccvm_solve_desc* desc, int32_t* info5);

int ccvm_abi_version(void);
const char* ccvm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CCVM_B200_H */
