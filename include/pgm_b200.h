/*
 * pgm_b200.h - C-ABI of the B200-native batched POGEMA step engine.
 *
 * The reference (CognitiveAISystems/pogema) is pure Python and has no FFI of
 * its own; the only mounted reference file is /root/reference/README.md:1-5
 * (a pointer to the GitHub repository).  Each entry point below therefore
 * cites the upstream Python symbol it replaces as `upstream <file> :: <symbol>`
 *  - no line numbers exist to cite.
 *
 * Conventions
 *   - every function returns 0 (PGM_OK) or a negative pgm_status; nothing
 *     throws across the ABI; pgm_last_error() gives the message of the last
 *     failure on the calling thread;
 *   - plain pointers and sizes only (no torch / C++ types);
 *   - pointers named *_dev are DEVICE pointers owned by the caller, pointers
 *     named *_host are host pointers owned by the caller; the engine owns only
 *     its handle and its internal state;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are stream-ordered and do not synchronise unless documented;
 *   - a handle is not thread-safe; handles on different devices are independent;
 *   - coordinates: x = row, y = column (upstream grid_config.py :: MOVES);
 *     the ABI speaks UNPADDED map coordinates, the engine stores padded ones.
 */
#ifndef PGM_B200_H_
#define PGM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGM_ABI_VERSION 1

typedef enum pgm_status {
  PGM_OK = 0,
  PGM_ERR_INVALID = -1,     /* bad argument / configuration                         */
  PGM_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the detail)  */
  PGM_ERR_UNSUPPORTED = -3, /* configuration outside what the kernels support       */
  PGM_ERR_OVERFLOW = -4,    /* upstream OverflowError: cannot place agents/targets  */
  PGM_ERR_STATE = -5,       /* call order violated (e.g. step before tasks exist)   */
  PGM_ERR_ACTION = -6       /* an action outside [0,5) was seen (sticky flag)       */
} pgm_status;

/* upstream grid_config.py :: GridConfig.collision_system */
enum { PGM_COLLISION_PRIORITY = 0, PGM_COLLISION_BLOCK_BOTH = 1, PGM_COLLISION_SOFT = 2 };
/* upstream grid_config.py :: GridConfig.on_target  (envs.py :: Pogema / PogemaCoopFinish / PogemaLifeLong) */
enum { PGM_ON_TARGET_FINISH = 0, PGM_ON_TARGET_NOTHING = 1, PGM_ON_TARGET_RESTART = 2 };
/* observation layouts written by pgm_step / pgm_observe */
enum {
  PGM_OBS_U8 = 0,  /* uint8 [N][A][3][D][D], values 0/1, channels obstacles/agents/target
                      (upstream envs.py :: _get_agents_obs; float32 upstream, same values)       */
  PGM_OBS_BITS = 1,/* uint32 [N][A][ceil(3*D*D/32)], bit k = element k of the uint8 layout       */
  PGM_OBS_F32 = 2, /* float32 [N][A][3][D][D], values 0.0/1.0 - the reference's own dtype         */
  PGM_OBS_F16 = 3  /* float16 [N][A][3][D][D], values 0.0/1.0 - for half-precision policies: the
                      kernel writes what `obs.half()` would, without the extra pass over the tensor */
};
/* pgm_get_state / pgm_state_ptr selectors */
enum {
  PGM_STATE_POSITIONS = 0, /* host copy: int32 [N][A][2] unpadded (x,y).  pgm_state_ptr: the raw
                              agent state array uint32 [N][A][2]: word 0 = (x+r) | active << 15 |
                              (y+r) << 16, word 1 = (tx+r) | (ty+r) << 16                         */
  PGM_STATE_TARGETS = 1,   /* int32 [N][A][2] unpadded (host copy only)                          */
  PGM_STATE_ACTIVE = 2,    /* uint8 [N][A]   upstream grid.py :: Grid.is_active (host copy only) */
  PGM_STATE_ELAPSED = 3,   /* int32 [N]      upstream MultiTimeLimit._elapsed_steps              */
  PGM_STATE_OBSTACLES = 4, /* uint8 [N][H][W] unpadded (host copy only)                          */
  PGM_STATE_WAS_ON_GOAL = 5,/* uint8 [N][A]  upstream envs.py :: Pogema.was_on_goal (last step)  */
  PGM_STATE_EPISODE_DONE = 6,/* uint8 [N]    all(terminated) or all(truncated) on the last step  */
  PGM_STATE_METRICS = 7,   /* int32 [N][4] raw counters of the last finished episode, from which
                              upstream wrappers/metrics.py values follow:
                              [0] sum of was_on_goal over the episode  (ISR = [0]/A, CSR = [0]==A,
                                  avg_throughput = [0]/max_episode_steps)
                              [1] sum of per-agent solve steps         (ep_length = [1]/A + 1)
                              [2] episode length in steps  [3] agents on goal at the last step     */
  PGM_STATE_SEEDS = 8,     /* uint64 [N] seed each instance's current task was built from (advances
                              under auto_reset=2)                                                  */
  PGM_STATE_SOLVE_COSTS = 9 /* int32 [N][A], on_target = nothing only: per-agent cost of the last finished
                              episode as upstream wrappers/metrics.py :: SumOfCostsAndMakespanMetric counts
                              it - the step at which the agent's final stay on its goal began, or the last
                              step (SoC = sum + A, makespan = max + 1).  pgm_state_ptr: the raw
                              int32 [N][A][2] array (running start of the stay | that cost)        */
};

typedef struct pgm_config {
  int32_t abi_version;       /* PGM_ABI_VERSION                                      */
  int32_t device;            /* CUDA device ordinal                                  */
  int32_t num_envs;          /* N instances held by this engine                      */
  int32_t num_agents;        /* A   upstream GridConfig.num_agents                   */
  int32_t height, width;     /* unpadded map size (GridConfig.size, or the map's)    */
  int32_t obs_radius;        /* r   upstream GridConfig.obs_radius, D = 2r+1         */
  int32_t max_episode_steps; /* upstream GridConfig.max_episode_steps                */
  int32_t collision_system;  /* PGM_COLLISION_*                                      */
  int32_t on_target;         /* PGM_ON_TARGET_*                                      */
  int32_t auto_reset;        /* 1: an instance whose episode ended is restored to its
                                initial state inside the same step and the returned
                                observation is the reset one (upstream
                                integrations/sample_factory.py :: AutoResetWrapper).
                                2: instead of restoring the same task, the instance is
                                REBUILT on the device from its next seed
                                (seed += reserved[0] or num_envs) right after the step -
                                upstream's behaviour with a changing seed (seed=None
                                draws a fresh map on every reset)                     */
  int32_t obs_format;        /* PGM_OBS_*                                            */
  int32_t team_threads;      /* threads cooperating on one instance; 0 = choose      */
  int32_t reserved[3];       /* [0]: seed stride of auto_reset=2 (0 = num_envs)         */
} pgm_config;

typedef struct pgm_engine pgm_engine;

const char* pgm_last_error(void);
int pgm_abi_version(void);

/* Lifetime. */
int pgm_create(const pgm_config* cfg, pgm_engine** out);
int pgm_destroy(pgm_engine* e);

/* Sizes the caller needs to allocate outputs. */
int64_t pgm_obs_bytes(const pgm_engine* e);            /* bytes of one full observation tensor     */
int64_t pgm_obs_instance_stride(const pgm_engine* e);  /* bytes between consecutive instances      */

/*
 * Task construction (replaces upstream grid.py :: Grid.__init__ /
 * GridLifeLong.__init__ and generator.py).  All three synchronise the stream
 * they upload on before returning.
 *
 * pgm_generate: instances [first, first+count) are generated from `seeds`
 *   exactly as upstream does with GridConfig(seed=seeds[k], size, density,
 *   num_agents): generator.py :: generate_obstacles (PCG64 binomial),
 *   bfs + generate_positions_and_targets_fast + placing, then
 *   grid.py :: add_artificial_border; for on_target=restart also
 *   generator.py :: get_components and the per-agent generators of
 *   envs.py :: PogemaLifeLong._initialize_grid.  If `map_host` is not NULL it is
 *   a uint8 [height][width] obstacle map shared by all instances
 *   (GridConfig.map) and only the placement is generated.
 *   Returns PGM_ERR_OVERFLOW if some instance cannot be placed (upstream
 *   raises OverflowError); *failed_index (optional) receives its index.
 * pgm_set_tasks: explicit obstacles / agents_xy / targets_xy (GridConfig.map +
 *   agents_xy + targets_xy).  obstacles_host: uint8 [count][H][W];
 *   agents_xy_host / targets_xy_host: int32 [count][A][2] unpadded.
 *   seeds (optional, may be NULL -> 0) feed the lifelong generators.
 */
int pgm_generate(pgm_engine* e, int32_t first, int32_t count, const uint64_t* seeds_host,
                 double density, const uint8_t* map_host, int32_t num_threads,
                 int32_t* failed_index, void* stream);
/* Same result as pgm_generate, built ON THE DEVICE (one warp per instance: PCG64 jump-ahead obstacle
 * draws, component labelling, shuffle, placing, border, lifelong tables) - three orders of magnitude
 * faster than the host path, for resets with new seeds.  Instances that need upstream's retry loop
 * are regenerated by the host generator (*num_host_fallbacks, optional, counts them). */
int pgm_generate_device(pgm_engine* e, int32_t first, int32_t count, const uint64_t* seeds_host,
                        double density, const uint8_t* map_host, int32_t* num_host_fallbacks,
                        void* stream);
int pgm_set_tasks(pgm_engine* e, int32_t first, int32_t count, const uint8_t* obstacles_host,
                  const int32_t* agents_xy_host, const int32_t* targets_xy_host,
                  const uint64_t* seeds_host, void* stream);

/*
 * Host-only generation of ONE instance (no CUDA call; the same generator
 * pgm_generate runs): upstream Grid.__init__ / GridLifeLong.__init__.
 *   obstacles_out [H][W] uint8 (after clearing cells under explicit starts/goals),
 *   agents_xy_out / targets_xy_out [A][2] int32 unpadded.
 *   lifelong != 0 additionally fills (if not NULL) rng_out [A][4] uint64
 *   {state_hi, state_lo, inc_hi, inc_lo} of each agent's PCG64 and comp_size_out [A].
 */
int pgm_generate_host(int32_t height, int32_t width, int32_t num_agents, int32_t obs_radius, double density,
                      int32_t lifelong, const uint8_t* map_host, uint64_t seed, uint8_t* obstacles_out,
                      int32_t* agents_xy_out, int32_t* targets_xy_out, uint64_t* rng_out,
                      int32_t* comp_size_out);

/*
 * upstream envs.py :: Pogema.reset (+ MultiTimeLimit.reset): restore every
 * instance to its initial state (same seed -> same map/task), elapsed = 0,
 * and, if obs_dev is not NULL, write the reset observations.
 */
int pgm_reset(pgm_engine* e, void* obs_dev, void* stream);

/* upstream envs.py :: Pogema._obs: write observations of the current state. */
int pgm_observe(pgm_engine* e, void* obs_dev, void* stream);

/*
 * upstream envs.py :: Pogema.step / PogemaLifeLong.step / PogemaCoopFinish.step
 * wrapped by wrappers/multi_time_limit.py :: MultiTimeLimit.step - ONE kernel
 * launch: move_agents (priority | block_both | soft), was_on_goal, rewards,
 * terminated, finish/restart bookkeeping, time limit, observations.
 *   actions_dev:   [N][A] integers in [0,5), element size action_itemsize (1,2,4,8 bytes,
 *                  little-endian; only the low byte is read)
 *   obs_dev:       observation tensor in the engine's PGM_OBS_* format (may be NULL to skip observations)
 *   rewards_dev:   float32 [N][A];  terminated_dev, truncated_dev: uint8 [N][A] (0/1)
 */
int pgm_step(pgm_engine* e, const void* actions_dev, int32_t action_itemsize, void* obs_dev,
             float* rewards_dev, uint8_t* terminated_dev, uint8_t* truncated_dev, void* stream);

/*
 * K consecutive steps in ONE launch for action tensors that are known in advance (open-loop
 * rollouts, random-policy data collection): exactly the results of K pgm_step calls, but every
 * instance runs its own timeline - no grid-wide barrier between steps, the map stays in shared
 * memory, the observation stores of one instance overlap the move phases of the others.
 *   actions_dev:  [K][N][A]        rewards/terminated/truncated_dev: [K][N][A]
 *   obs_dev:      [obs_ring][N][A]... step k writes slot k % obs_ring (NULL skips observations)
 */
int pgm_step_many(pgm_engine* e, int32_t num_steps, const void* actions_dev, int32_t action_itemsize,
                  void* obs_dev, int32_t obs_ring, float* rewards_dev, uint8_t* terminated_dev,
                  uint8_t* truncated_dev, void* stream);

/*
 * Same step with HOST buffers: copies actions host->device, launches the step,
 * copies obs / rewards / flags device->host and synchronises the stream.  Host
 * buffers may be pageable or pinned (pinned is faster).  obs_host may be NULL.
 */
int pgm_step_host(pgm_engine* e, const void* actions_host, int32_t action_itemsize, void* obs_host,
                  float* rewards_host, uint8_t* terminated_host, uint8_t* truncated_host,
                  void* stream);

/*
 * pgm_step_host plus the two per-agent flags upstream keeps beside the step results, in the same
 * synchronisation: active_host uint8 [N][A] = upstream grid.py :: Grid.is_active after the step,
 * was_on_goal_host uint8 [N][A] = upstream envs.py :: Pogema.was_on_goal.  Either may be NULL.
 * (The list API needs both every step for `infos`; asking for them here instead of through two
 * pgm_get_state calls takes a single-instance list-API step from 124 us to 48 us.)
 */
int pgm_step_host_ex(pgm_engine* e, const void* actions_host, int32_t action_itemsize, void* obs_host,
                     float* rewards_host, uint8_t* terminated_host, uint8_t* truncated_host,
                     uint8_t* active_host, uint8_t* was_on_goal_host, void* stream);

/* upstream Pogema._obs with a HOST destination buffer (copies device->host, synchronises). */
int pgm_observe_host(pgm_engine* e, void* obs_host, void* stream);

/*
 * How pgm_step_host / pgm_observe_host bring PGM_OBS_U8 / PGM_OBS_F16 / PGM_OBS_F32 observations to the host.
 *   mode 0 (plain):  the kernel writes the final tensor, the copy engine moves all of it (PCIe-bound:
 *                    3*D*D bytes, or 2x / 4x that, per agent).
 *   mode 1 (packed): the kernel writes the observation BIT STREAM (1 bit per element, every bit still
 *                    computed on the GPU), the copy engine moves 1/8 (1/16, 1/32) of the bytes in chunks into
 *                    pinned staging owned by the engine, and `num_threads` host threads (0 = all cores,
 *                    at most 32) widen bit k to element k of obs_host with non-temporal stores while
 *                    later chunks are still on the bus.  obs_host receives exactly the bytes of mode 0.
 *   mode -1 (auto, the default): packed when the tensor is >= 4 MB (env PGM_HOST_TRANSPORT=0/1 overrides).
 */
int pgm_set_host_transport(pgm_engine* e, int32_t mode, int32_t num_threads);
/* out[0] 1 if the next host call uses the packed transport, [1] host threads of the pool (0 = not started),
 * [2]/[3] host->device / device->host bytes moved by the last pgm_step_host, [4] host ISA of the widening
 * loop (0 scalar, 1 AVX2, 2 AVX-512BW), [5..9] microseconds from the entry of the last packed
 * pgm_step_host to: all work enqueued, first chunk on the host, last chunk on the host, widening done,
 * stream idle (= return). */
int pgm_host_transport_info(const pgm_engine* e, int64_t* out, int32_t n);
/* The widening loop of the packed transport on its own (no CUDA call; used by the CPU tests): bit k of
 * the little-endian word stream src_host -> element k of dst_host (elem_size 1: uint8 0/1, 2: float16, 4: float32 0.0/1.0). */
int pgm_expand_bits_host(const uint32_t* src_host, int64_t nbits, void* dst_host, int32_t elem_size);
/* Measurement aid (no CUDA call): GB/s of a plain non-temporal fill of dst_host[0..bytes) by num_threads host
 * threads, `reps` passes - the DRAM write ceiling of this host, i.e. the bound of the packed transport's
 * widening loop (bench.py reports it beside `e2e`).  Returns a negative value on bad arguments. */
double pgm_host_fill_gbps(void* dst_host, int64_t bytes, int32_t num_threads, int32_t reps);

/* State access (debugging, env.grid accessors, checkpoint/resume). */
int pgm_get_state(pgm_engine* e, int32_t what, void* dst_host, int64_t dst_bytes, void* stream);
void* pgm_state_ptr(pgm_engine* e, int32_t what); /* device pointer of the raw array, or NULL */

/* Checkpoint / resume of the complete mutable state: positions, targets, active flags, elapsed steps, lifelong
 * generators, metric counters, current task seeds; with auto_reset == 2 (tasks rebuilt from new seeds every
 * episode) also the tasks themselves (obstacle bitmaps, initial states, lifelong tables).  The blob starts with a
 * 64-byte header (magic, ABI, shape, modes): pgm_checkpoint_load rejects a blob of another engine shape or mode
 * and, when the tasks are not part of the blob, a blob taken from an engine built from other seeds. */
int64_t pgm_checkpoint_bytes(const pgm_engine* e);
int pgm_checkpoint_save(pgm_engine* e, void* dst_host, int64_t dst_bytes, void* stream);
int pgm_checkpoint_load(pgm_engine* e, const void* src_host, int64_t src_bytes, void* stream);

/* Sticky device-side error flag (PGM_ERR_ACTION ...); reading it synchronises. */
int pgm_check_errors(pgm_engine* e, void* stream);

/* Development aid: if dev_ptr is not NULL every later launch stores clock64() stamps at its phase
 * boundaries into int64 [N][16] at dev_ptr (caller-owned device memory); NULL switches it off.  The stamp code is
 * compiled only into the timeline build of the library (`make timeline`, -DPGM_TIMELINE); in the product build the
 * call is accepted and the kernels write nothing. */
int pgm_set_debug_buffer(pgm_engine* e, void* dev_ptr);

/* Number of kernel launches issued by this engine so far (bench `gpu_launches`). */
int64_t pgm_launch_count(const pgm_engine* e);

/* Kernel plan actually chosen (for DESIGN.md / bench config): fills up to n ints:
 * [0] team_threads [1] teams_per_cta [2] cta_threads [3] smem_bytes_per_cta
 * [4] grid [5] agents_per_obs_batch [6] occupancy structure (0 dense cell grid, 1 tile buckets)
 * [7] 1 if step launches use the register-resident kernel for the common shapes (pgm_fast.cuh), then its
 * [8] team_threads [9] agents per thread [10] teams_per_cta [11] smem_bytes_per_cta [12] grid */
int pgm_plan(const pgm_engine* e, int32_t* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* PGM_B200_H_ */
