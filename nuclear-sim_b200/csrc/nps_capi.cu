// nps_b200: CUDA kernels (sm_100a) and the C ABI declared in include/nps_b200.h.
//
// Execution model: plants are independent, so the step path is pure data parallelism — one
// thread advances one plant through k fused substeps.  Plant state lives in HBM as an FP64
// structure-of-arrays slab (field-major), so every load/store of a field by a warp is one fully
// coalesced 256-byte transaction.  There is no dense contraction anywhere on this path: no tensor
// cores, the bounds are the FP64 pipe and HBM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../../include/nps_b200.h"
#include "plant/plant_step.h"
#include "plant/maintenance.h"
#include "plant/rng.h"
#include "fields_gen.inc"

using namespace nps;

static thread_local std::string g_last_error;
static int fail(const char* what, cudaError_t e = cudaSuccess) {
    g_last_error = what;
    if (e != cudaSuccess) { g_last_error += ": "; g_last_error += cudaGetErrorString(e); }
    return -1;
}
#define NPS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(#call, e__); } while (0)

// every entry point that launches or copies runs on the handle's device and leaves the caller's current device alone
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

constexpr int kNState = (int)(sizeof(PlantState) / sizeof(double));
constexpr int kNParams = (int)(sizeof(PlantParams) / sizeof(double));
static_assert(kNState == NPS_GEN_N_STATE, "state.h and fields_gen.inc disagree; rerun the build");
static_assert(kNParams == NPS_GEN_N_PARAMS, "state.h and fields_gen.inc disagree; rerun the build");

// device-side noise (csrc/plant/rng.h): used when the caller passes no noise array and nps_set_device_rng enabled it
struct RngConfig { uint64_t seed; uint64_t plant_offset; uint64_t step0; int enabled; int pad; };
struct Threshold { int field; int cmp; double value; double cooldown; int row; int pad; };
__device__ __forceinline__ bool threshold_compare(int cmp, double v, double x) {   // _check_threshold_condition: state_manager.py:1412-1442
    switch (cmp) {
        case 0: return v > x;
        case 1: return v < x;
        case 2: return v >= x;
        case 3: return v <= x;
        case 4: return fabs(v - x) < 1e-3;
        case 5: return fabs(v - x) >= 1e-3;
        default: return false;
    }
}

struct nps_handle {
    int64_t n = 0;
    int device = 0;
    PlantParams params;
    // staging buffers for nps_step_host
    int8_t* d_action = nullptr; double* d_mag = nullptr; double* d_noise = nullptr; double* d_setpoint = nullptr;
    double* d_obs = nullptr; double* d_reward = nullptr; uint8_t* d_done = nullptr;
    int staged_k = 0;
    Threshold* d_thresholds = nullptr; int n_thresholds = 0; int n_live_thresholds = 0;
    int32_t* d_word_start = nullptr;   // live rows are sorted by table row: [word w] = d_thresholds[d_word_start[w] .. d_word_start[w + 1])
    int32_t* d_logged = nullptr; int n_logged = 0;
    int32_t* d_gather_fields = nullptr; double* d_gather_out = nullptr; int gather_cap = 0;
    // nps_apply_maintenance: request / group / status buffers grow geometrically and are reused (no cudaMalloc per call);
    // the host mirrors are pinned so the copies are truly asynchronous on the caller's stream
    struct MaintBuf { void* d_req = nullptr; int32_t* d_start = nullptr; int32_t* d_status = nullptr;
                      void* h_req = nullptr; int32_t* h_start = nullptr; int32_t* h_status = nullptr; int cap = 0; } mb;
    // nps_step_host_async: NPS_PIPE_DEPTH staging sets, so the host->device copy of a launch is issued several
    // kernels ahead of its use (a sporadically slow PCIe transfer then costs nothing) and overlaps the running kernel
    struct Pipe {
        int8_t* d_action = nullptr; double* d_mag = nullptr; double* d_noise = nullptr; double* d_setpoint = nullptr;
        double* d_obs = nullptr; double* d_reward = nullptr; uint8_t* d_done = nullptr;
        cudaEvent_t in_done = nullptr, kernel_done = nullptr, out_done = nullptr;
        cudaEvent_t t_in0 = nullptr, t_k0 = nullptr;   // NPS_PIPE_TRACE=1: timed events around the input copy and the kernel
    } pipe[NPS_PIPE_DEPTH];
    bool pipe_trace = false;
    cudaStream_t copy_stream = nullptr, out_stream = nullptr;   // host->device and device->host on separate streams
    int pipe_k = 0; int64_t pipe_count = 0;
    RngConfig rng = {0, 0, 0, 0, 0};
    int n_sms = 148; bool log_row_tile_only = false;
    // from this many plants on: 448 threads x 128 registers per SM.  Up to 37,888 uncapped one-warp blocks are one wave
    // and 4-8 % faster (profiles/r02_mid_batch.txt); NPS_LARGE_BATCH overrides (tuning)
    int64_t large_batch = 148 * 8 * 32;
    int small_shape = 0;   // batches up to n_sms x 4 x 32 plants: 0 split kernel (two threads per plant), 1 one thread per plant (NPS_SMALL_SHAPE=1)   // NPS_LOG_ROW_TILE=1 forces the shared-memory tile kernel
};

// ------------------------------------------------------------------------------------------------
// step kernel: thread-per-plant, state register/local resident across k substeps
// ------------------------------------------------------------------------------------------------
// Launch shapes:
//   large batches  448 threads x 1 block/SM, 128 registers: 65,536 plants are ONE wave on 148 SMs
//                  (measured 2 % faster than 64 x 7 at 65,536 plants, profiles/r01_tuning_variants.txt (7))
//   small batches  one warp per block, no register cap (ptxas takes what it needs up to 255): below ~33 K plants an SM
//                  holds at most 7 warps, so registers are free and the time is one warp's own dependency chain - a
//                  wider register file per thread means fewer spills on that chain (profiles/r01_tuning_variants.txt (21))
// NPS_STEP_BLOCK / NPS_STEP_MINBLOCKS pin one shape for tuning builds.
#ifndef NPS_COPY_UNROLL
#define NPS_COPY_UNROLL 8   /* loads in flight per thread while the slab is copied in / out; 2, 4, 24 and 48 all
                               measured slower on B200 (profiles/r01_tuning_variants.txt) */
#endif
#define NPS_STR_(x) #x
#define NPS_PRAGMA_UNROLL(n) _Pragma(NPS_STR_(unroll n))

// ------------------------------------------------------------------------------------------------
// in-launch monitoring: what the reference does after the physics of EVERY step (sim.py:209-223,256) evaluated after
// every fused substep, so a K-substep launch reports the same (event, step) pairs as K single-step launches:
//   * threshold rows (StateManager._check_maintenance_thresholds: state_manager.py:1307-1369) with their cooldown
//     stamps; violations are appended to a device event list (warp ballot -> one atomic per warp per row),
//   * first step a watched flag field became non-zero (latched trips, FSM states: SURVEY 8 a-events e3-e12),
//   * first scram step / first NaN-reset step / sticky status bits, per-substep reward and done.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSharedRows = 128;
struct MonitorArgs {
    int enabled, n_live, skip_last_check, n_watch;
    const Threshold* live; double* last_fired;
    nps_event* events; uint32_t* n_events; uint32_t event_cap; uint32_t pad;
    const int32_t* watch_fields; int32_t* watch_step;
    int32_t* first_scram_step; int32_t* first_nan_reset_step; uint32_t* status;
    double* reward_k; uint8_t* done_k;
    int64_t step0;
};
struct StepArgs {
    double* slab; const int8_t* action; const double* magnitude; const double* noise; const double* setpoint;
    double* obs; double* reward; uint8_t* done;
    int64_t n; int k_substeps; int pad;
    RngConfig rng; MonitorArgs mon;
};

__device__ __forceinline__ double derived_from_state(const PlantState& st, int code) {
    const int k = -code - 2, kind = k >> 2, unit = k & 3;
    if (kind == 0) {   // sum_wear_level: feedwater/pump_lubrication.py:1585-1596
        const double* w = st.fw.pump[unit].lub.component_wear;
        return w[0] + py_max3(w[FWL_MOTOR_BRG], w[FWL_PUMP_BRG], w[FWL_THRUST_BRG]) + w[FWL_SEALS];
    }
    return NAN;
}


// Which half of a split step (plant_step_source / plant_step_sink, csrc/plant/plant_step.h) owns state field f: the sink
// half owns the turbine and condenser records, the six SecondaryState fields, sim.last_load_factor and the heat-flow
// report fields it writes; the source half owns everything else.  A half's frame holds valid values for its own fields only.
#define NPS_FIELD_OF(member) ((int)(offsetof(PlantState, member) / sizeof(double)))
__device__ __forceinline__ bool sink_owns(int f) {
    constexpr int lo = NPS_FIELD_OF(turb), hi = NPS_FIELD_OF(ph);     // [turb, cond] are adjacent, ph follows
    static_assert(NPS_FIELD_OF(cond) > NPS_FIELD_OF(turb) && NPS_FIELD_OF(ph) > NPS_FIELD_OF(cond), "PlantState order changed");
    if (f >= lo && f < hi) return true;
    return f == NPS_FIELD_OF(sec.total_system_heat_rejection) || f == NPS_FIELD_OF(sec.power_reduction_factor) ||
           f == NPS_FIELD_OF(sec.electrical_power_output) || f == NPS_FIELD_OF(sec.thermal_efficiency) ||
           f == NPS_FIELD_OF(sec.heat_rate_kj_kwh) || f == NPS_FIELD_OF(sec.condenser_pressure) ||
           f == NPS_FIELD_OF(sim.last_load_factor) ||
           (f >= NPS_FIELD_OF(rep.hf_steam_enthalpy_flow) && f <= NPS_FIELD_OF(rep.hf_energy_balance_percent));
}

// role: -1 whole plant (one thread per plant), 0 source half, 1 sink half (evaluates only the rows / watched fields it owns;
// `now` is the plant clock after the step, which a sink half gets from its source half)
// The monitor reads ~100 scattered fields of the frame once per substep.  NPS_MON_LD selects the cache operator of those
// loads: 0 plain, 1 ld.local.cs (streaming: evict first), 2 ld.local.lu (last use) - so that they do not push the lines
// the next substep starts with out of L1 (profiles/r02_monitor_cost.txt).
#ifndef NPS_MON_LD
#define NPS_MON_LD 0
#endif
__device__ __forceinline__ double mon_load(const double* __restrict__ sv, int f) {
#if NPS_MON_LD == 0
    return sv[f];
#else
    double v;
    unsigned long long a;
    asm volatile("cvta.to.local.u64 %0, %1;" : "=l"(a) : "l"(sv + f));
#if NPS_MON_LD == 1
    asm volatile("ld.local.cs.f64 %0, [%1];" : "=d"(v) : "l"(a));
#else
    asm volatile("ld.local.lu.f64 %0, [%1];" : "=d"(v) : "l"(a));
#endif
    return v;
#endif
}

__device__ __noinline__ void monitor_substep(const PlantState& st, const MonitorArgs& mon_ref, const Threshold* __restrict__ rows,
                                             int64_t n, int64_t p, int k, bool last, unsigned warp_mask, unsigned step_status,
                                             unsigned& seen_watch_ref, int role = -1, double now_in = 0.0) {
    const MonitorArgs mon = mon_ref;          // by value: the stores below cannot alias the argument record
    unsigned seen_watch = seen_watch_ref;
    const double* __restrict__ sv = reinterpret_cast<const double*>(&st);
    const int32_t step = (int32_t)(mon.step0 + k);
    if (step_status) {
        if (mon.status) mon.status[p] |= step_status;
        if ((step_status & kStatusScram) && mon.first_scram_step && mon.first_scram_step[p] < 0) mon.first_scram_step[p] = step;
        if ((step_status & kStatusNanReset) && mon.first_nan_reset_step && mon.first_nan_reset_step[p] < 0) mon.first_nan_reset_step[p] = step;
    }
    // Watched flags and threshold values are first touches of this substep's freshly written state, i.e. cache misses:
    // every group of kMonChunk values is loaded with NO control flow between the loads (rows this thread does not own
    // read field 0 and are masked afterwards), so a group costs one memory round trip instead of eight
    // (profiles/r02_monitor_cost.txt).
#ifndef NPS_MON_CHUNK
#define NPS_MON_CHUNK 8
#endif
    constexpr int kMonChunk = NPS_MON_CHUNK;
    for (int w0 = 0; w0 < mon.n_watch; w0 += kMonChunk) {
        double wv[kMonChunk];
        int wf[kMonChunk];
#pragma unroll
        for (int j = 0; j < kMonChunk; ++j) wf[j] = (w0 + j < mon.n_watch) ? __ldg(mon.watch_fields + w0 + j) : 0;
#pragma unroll
        for (int j = 0; j < kMonChunk; ++j) wv[j] = mon_load(sv, wf[j]);
#pragma unroll
        for (int j = 0; j < kMonChunk; ++j) {
            const int w = w0 + j;
            const bool own = (w < mon.n_watch) && (role < 0 || sink_owns(wf[j]) == (role == 1));
            if (own && !((seen_watch >> w) & 1u) && wv[j] != 0.0) {
                mon.watch_step[(int64_t)w * n + p] = step;
                seen_watch |= 1u << w;
            }
        }
    }
    seen_watch_ref = seen_watch;
    if (mon.last_fired && !(last && mon.skip_last_check)) {
        const double now = (role == 1) ? now_in : st.sim.time_minutes;
        const unsigned lane = threadIdx.x & 31u;
        // the warp votes ONCE per group, and only a group in which some lane fired goes through the per-row ballots
        // that build the event list
        for (int j0 = 0; j0 < mon.n_live; j0 += kMonChunk) {
            double v[kMonChunk];
            int fld[kMonChunk];
            unsigned mine = 0, hit = 0;
#pragma unroll
            for (int j = 0; j < kMonChunk; ++j) {
                const int f = (j0 + j < mon.n_live) ? rows[j0 + j].field : -1;      // -1: padding
                fld[j] = f;
                const bool own = (f != -1) && (role < 0 || ((f >= 0 && sink_owns(f)) == (role == 1)));   // derived rows: source half
                mine |= (own ? 1u : 0u) << j;
            }
#pragma unroll
            for (int j = 0; j < kMonChunk; ++j) v[j] = mon_load(sv, (((mine >> j) & 1u) && fld[j] >= 0) ? fld[j] : 0);
#pragma unroll
            for (int j = 0; j < kMonChunk; ++j) {
                if (!((mine >> j) & 1u)) continue;
                if (fld[j] < -1) v[j] = derived_from_state(st, fld[j]);
                const double x = rows[j0 + j].value;
                const int c = rows[j0 + j].cmp;
                const double d = fabs(v[j] - x);
                const bool fire = (c == 0 && v[j] > x) || (c == 1 && v[j] < x) || (c == 2 && v[j] >= x) || (c == 3 && v[j] <= x) ||
                                  (c == 4 && d < 1e-3) || (c == 5 && d >= 1e-3);      // _check_threshold_condition
                hit |= (fire ? 1u : 0u) << j;
            }
            if (!__ballot_sync(warp_mask, hit != 0)) continue;
#pragma unroll
            for (int j = 0; j < kMonChunk; ++j) {
                if (j0 + j >= mon.n_live) break;
                const Threshold th = rows[j0 + j];
                bool fire = (hit >> j) & 1u;
                if (fire) {   // _is_threshold_in_cooldown (state_manager.py:1267-1305) only matters for rows that would fire
                    double* stamp = mon.last_fired + (int64_t)th.row * n + p;
                    if ((now - *stamp) < th.cooldown) fire = false;
                    else *stamp = now;      // _record_threshold_violation_time
                }
                const unsigned m = __ballot_sync(warp_mask, fire);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    uint32_t base = 0;
                    if ((int)lane == leader) base = atomicAdd(mon.n_events, (uint32_t)__popc(m));
                    base = __shfl_sync(warp_mask, base, leader);
                    if (fire) {
                        const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
                        if (slot < mon.event_cap) mon.events[slot] = nps_event{(int32_t)p, th.row, step, 0, v[j], now};
                    }
                }
            }
        }
    }
}

// 448 threads get 128 registers: warps are allocated in fours, so the block counts as 16 warps and 136 or 144 registers
// per thread (which 14 x 32 threads would fit) fail to launch (checked with __maxnreg__, GPU call 21).
template <int BLOCK, int MINBLOCKS>
__global__ void __launch_bounds__(BLOCK, MINBLOCKS)
nps_step_kernel(const __grid_constant__ PlantParams prm, const __grid_constant__ StepArgs a) {
    __shared__ Threshold s_rows[kMaxSharedRows];
    const MonitorArgs& mon = a.mon;
    const Threshold* rows = mon.live;
    if (mon.enabled && mon.n_live > 0 && mon.n_live <= kMaxSharedRows) {
        for (int j = threadIdx.x; j < mon.n_live; j += BLOCK) s_rows[j] = mon.live[j];
        __syncthreads();
        rows = s_rows;
    }
    const int64_t n = a.n;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned warp_mask = __ballot_sync(0xffffffffu, p < n);
    if (p >= n) return;
    double* __restrict__ slab = a.slab;
    PlantState st;
    double* sv = reinterpret_cast<double*>(&st);
NPS_PRAGMA_UNROLL(NPS_COPY_UNROLL)
    for (int f = 0; f < kNState; ++f) sv[f] = slab[(int64_t)f * n + p];
    bool scrammed = false;
    unsigned seen_watch = 0;
    if (mon.enabled)
        for (int w = 0; w < mon.n_watch; ++w) if (mon.watch_step[(int64_t)w * n + p] >= 0) seen_watch |= 1u << w;
    const int k_substeps = a.k_substeps;
    for (int k = 0; k < k_substeps; ++k) {
        StepInput in;
        in.action = a.action ? (int)a.action[(int64_t)k * n + p] : (int)ACT_NO_ACTION;
        in.magnitude = a.magnitude ? a.magnitude[(int64_t)k * n + p] : 1.0;
        if (a.noise) {
            const double* z = a.noise + (int64_t)k * NPS_NOISE_PER_STEP * n + p;
            in.z_heat = z[0]; in.z_ph = z[n]; in.u_ph[0] = z[2 * n]; in.u_ph[1] = z[3 * n]; in.u_ph[2] = z[4 * n];
        } else if (a.rng.enabled) {
            const StepDraws d = plant_step_draws(a.rng.seed, a.rng.plant_offset + (uint64_t)p, a.rng.step0 + (uint64_t)k);
            in.z_heat = d.z_heat; in.z_ph = d.z_ph; in.u_ph[0] = d.u_ph[0]; in.u_ph[1] = d.u_ph[1]; in.u_ph[2] = d.u_ph[2];
        } else {
            in.z_heat = 0.0; in.z_ph = 0.0; in.u_ph[0] = 1.0; in.u_ph[1] = 1.0; in.u_ph[2] = 1.0;
        }
        in.power_setpoint = a.setpoint ? a.setpoint[(int64_t)k * n + p] : NAN;
        in.emit_outputs = (k == k_substeps - 1);
        plant_step(st, prm, in);
        const bool scram_now = is_true(st.pri.scram_activated);
        scrammed |= scram_now;
        if (mon.enabled) {
            monitor_substep(st, mon, rows, n, p, k, k == k_substeps - 1, warp_mask, in.status, seen_watch);
            if (mon.reward_k) mon.reward_k[(int64_t)k * n + p] = plant_reward(st, prm);
            if (mon.done_k) mon.done_k[(int64_t)k * n + p] = scram_now ? 1 : 0;
        }
    }
NPS_PRAGMA_UNROLL(NPS_COPY_UNROLL)
    for (int f = 0; f < kNState; ++f) slab[(int64_t)f * n + p] = sv[f];
    if (a.obs) {
        struct ObsOut { double* o; int64_t n; int64_t p; struct Ref { double* a; NPS_HD void operator=(double v) { *a = v; } };
                        __device__ Ref operator[](int i) { return Ref{o + (int64_t)i * n + p}; } } out{a.obs, n, p};
        plant_observe(st, prm, out);
    }
    if (a.reward) a.reward[p] = plant_reward(st, prm);
    if (a.done) a.done[p] = scrammed ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// split launch shape for small batches: TWO threads per plant, pipelined by one substep
// ------------------------------------------------------------------------------------------------
// Below ~33 K plants a B200 has fewer than 8 warps per SM and the time of a launch is ONE warp's dependency chain
// (DESIGN.md 5).  The turbine + condenser are pure sinks of a step's dataflow (csrc/plant/secondary.h SecHandoff), so a
// plant is advanced by a source thread (primary, feedwater, steam generators, chemistry, clock) and a sink thread
// (turbine, condenser, energy bookkeeping): while the source half computes substep k, the sink half computes substep
// k - 1 from the handoff the source left in shared memory.  Same functions, same arithmetic, same order inside each
// half -> bit-identical state; the chain per substep drops from source + sink to max(source, sink).
// Block = 64 threads = 32 plants: warp 0 source halves, warp 1 sink halves, one __syncthreads per substep.
constexpr int kSplitPlants = 32;
struct SplitShared {
    SecHandoff hand[2][kSplitPlants];
    double now[2][kSplitPlants];            // plant clock after the substep (the sink half stamps its events with it)
    double a2b[2][6][kSplitPlants];         // power level, fuel temperature, coolant pressure, scram status, load demand, SG pressure: reward
    double b2a[3][kSplitPlants];            // electrical output, thermal efficiency, condenser pressure: final observation / reward
};

// No register cap (224 registers, 4 blocks per SM): one wave up to 148 x 4 x 32 = 18,944 plants.  A 128-register
// variant that keeps 33,152 plants in one wave was measured and dropped: at 32,768 plants it reaches 6.9e7 plant-steps/s
// against 1.0e8 for one uncapped thread per plant (profiles/r02_bench_n1.json small_batch), so larger batches use that.
// Coupling the halves through a ring of 2-4 handoff slots with producer/consumer counters instead of the block barrier,
// and giving L1 more of the SM (64 KB instead of 100 KB of shared memory), were both measured: no change outside run-to-run
// noise at 1,024-18,944 plants, deeper rings lose L1 and are slower (profiles/r02_split_ring.txt).
__global__ void __launch_bounds__(2 * kSplitPlants, 1)
nps_step_split_kernel(const __grid_constant__ PlantParams prm, const __grid_constant__ StepArgs a) {
    __shared__ Threshold s_rows[kMaxSharedRows];
    __shared__ SplitShared sh;
    const MonitorArgs& mon = a.mon;
    const Threshold* rows = mon.live;
    if (mon.enabled && mon.n_live > 0 && mon.n_live <= kMaxSharedRows) {
        for (int j = threadIdx.x; j < mon.n_live; j += 2 * kSplitPlants) s_rows[j] = mon.live[j];
        rows = s_rows;
    }
    __syncthreads();
    const int role = threadIdx.x >> 5;                 // 0 source half, 1 sink half
    const int lane = threadIdx.x & 31;
    const int64_t n = a.n;
    const int64_t p = (int64_t)blockIdx.x * kSplitPlants + lane;
    const bool valid = p < n;
    const unsigned warp_mask = __ballot_sync(0xffffffffu, valid);
    double* __restrict__ slab = a.slab;
    PlantState st;
    double* sv = reinterpret_cast<double*>(&st);
    if (valid) {
NPS_PRAGMA_UNROLL(NPS_COPY_UNROLL)
        for (int f = 0; f < kNState; ++f) if (sink_owns(f) == (role == 1)) sv[f] = slab[(int64_t)f * n + p];
    }
    bool scrammed = false;
    unsigned seen_watch = 0;
    if (mon.enabled && valid)
        for (int w = 0; w < mon.n_watch; ++w) if (mon.watch_step[(int64_t)w * n + p] >= 0) seen_watch |= 1u << w;
    const int K = a.k_substeps;
#if defined(NPS_SPLIT_TIMING)
    long long busy = 0;      // tuning build: cycles each half spends working (not waiting at the barrier)
#endif
    for (int k = 0; k <= K; ++k) {
#if defined(NPS_SPLIT_TIMING)
        const long long c0 = clock64();
#endif
        if (role == 0 && k < K && valid) {
            StepInput in;
            in.action = a.action ? (int)a.action[(int64_t)k * n + p] : (int)ACT_NO_ACTION;
            in.magnitude = a.magnitude ? a.magnitude[(int64_t)k * n + p] : 1.0;
            if (a.noise) {
                const double* z = a.noise + (int64_t)k * NPS_NOISE_PER_STEP * n + p;
                in.z_heat = z[0]; in.z_ph = z[n]; in.u_ph[0] = z[2 * n]; in.u_ph[1] = z[3 * n]; in.u_ph[2] = z[4 * n];
            } else if (a.rng.enabled) {
                const StepDraws d = plant_step_draws(a.rng.seed, a.rng.plant_offset + (uint64_t)p, a.rng.step0 + (uint64_t)k);
                in.z_heat = d.z_heat; in.z_ph = d.z_ph; in.u_ph[0] = d.u_ph[0]; in.u_ph[1] = d.u_ph[1]; in.u_ph[2] = d.u_ph[2];
            } else {
                in.z_heat = 0.0; in.z_ph = 0.0; in.u_ph[0] = 1.0; in.u_ph[1] = 1.0; in.u_ph[2] = 1.0;
            }
            in.power_setpoint = a.setpoint ? a.setpoint[(int64_t)k * n + p] : NAN;
            in.emit_outputs = (k == K - 1);
            SecHandoff h;
            h.emit_outputs = in.emit_outputs;
            plant_step_source(st, prm, in, h);
            sh.hand[k & 1][lane] = h;
            sh.now[k & 1][lane] = st.sim.time_minutes;
            const bool scram_now = is_true(st.pri.scram_activated);
            scrammed |= scram_now;
            if (mon.enabled) {
                monitor_substep(st, mon, rows, n, p, k, k == K - 1, warp_mask, in.status, seen_watch, 0);
                if (mon.done_k) mon.done_k[(int64_t)k * n + p] = scram_now ? 1 : 0;
                if (mon.reward_k) {
                    sh.a2b[k & 1][0][lane] = st.pri.power_level; sh.a2b[k & 1][1][lane] = st.pri.fuel_temperature;
                    sh.a2b[k & 1][2][lane] = st.pri.coolant_pressure; sh.a2b[k & 1][3][lane] = st.pri.scram_status;
                    sh.a2b[k & 1][4][lane] = st.sim.load_demand; sh.a2b[k & 1][5][lane] = st.sec.sg_avg_pressure;
                }
            }
        } else if (role == 1 && k >= 1 && valid) {
            const int j = k - 1, b = j & 1;
            const SecHandoff h = sh.hand[b][lane];
            plant_step_sink(st, prm, h);
            if (mon.enabled) {
                monitor_substep(st, mon, rows, n, p, j, j == K - 1, warp_mask, 0u, seen_watch, 1, sh.now[b][lane]);
                if (mon.reward_k) {   // calculate_reward needs both halves: the six source-side values travel with the handoff
                    st.pri.power_level = sh.a2b[b][0][lane]; st.pri.fuel_temperature = sh.a2b[b][1][lane];
                    st.pri.coolant_pressure = sh.a2b[b][2][lane]; st.pri.scram_status = sh.a2b[b][3][lane];
                    st.sim.load_demand = sh.a2b[b][4][lane]; st.sec.sg_avg_pressure = sh.a2b[b][5][lane];
                    mon.reward_k[(int64_t)j * n + p] = plant_reward(st, prm);
                }
            }
        }
#if defined(NPS_SPLIT_TIMING)
        busy += clock64() - c0;
#endif
        __syncthreads();
    }
#if defined(NPS_SPLIT_TIMING)
    if (blockIdx.x == 0 && lane == 0) printf("[split timing] role %d busy cycles per substep %lld (K = %d)\n", role, busy / K, K);
#endif
    if (valid) {
NPS_PRAGMA_UNROLL(NPS_COPY_UNROLL)
        for (int f = 0; f < kNState; ++f) if (sink_owns(f) == (role == 1)) slab[(int64_t)f * n + p] = sv[f];
    }
    if (role == 1 && valid) {
        sh.b2a[0][lane] = st.sec.electrical_power_output; sh.b2a[1][lane] = st.sec.thermal_efficiency;
        sh.b2a[2][lane] = st.sec.condenser_pressure;
    }
    __syncthreads();
    if (role == 0 && valid) {
        st.sec.electrical_power_output = sh.b2a[0][lane]; st.sec.thermal_efficiency = sh.b2a[1][lane];
        st.sec.condenser_pressure = sh.b2a[2][lane];
        if (a.obs) {
            struct ObsOut { double* o; int64_t n; int64_t p; struct Ref { double* a; NPS_HD void operator=(double v) { *a = v; } };
                            __device__ Ref operator[](int i) { return Ref{o + (int64_t)i * n + p}; } } out{a.obs, n, p};
            plant_observe(st, prm, out);
        }
        if (a.reward) a.reward[p] = plant_reward(st, prm);
        if (a.done) a.done[p] = scrammed ? 1 : 0;
    }
}

template <class T>
__device__ __forceinline__ void load_member(T& dst, const PlantState& base, const double* __restrict__ slab, int64_t n, int64_t p) {
    const int f0 = (int)(reinterpret_cast<const double*>(&dst) - reinterpret_cast<const double*>(&base));
    double* d = reinterpret_cast<double*>(&dst);
    for (int f = 0; f < (int)(sizeof(T) / sizeof(double)); ++f) d[f] = slab[(int64_t)(f0 + f) * n + p];
}

__global__ void nps_observe_kernel(const double* __restrict__ slab, const __grid_constant__ PlantParams prm, int64_t n,
                                   double* __restrict__ obs, double* __restrict__ reward) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    PlantState st;   // only the members read by plant_observe / plant_reward are loaded
    load_member(st.pri, st, slab, n, p);
    load_member(st.sim, st, slab, n, p);
    load_member(st.sec, st, slab, n, p);
    load_member(st.fw.total_flow_rate, st, slab, n, p);
    load_member(st.fw.total_power_consumption, st, slab, n, p);
    load_member(st.fw.system_availability, st, slab, n, p);
    if (obs) {
        struct ObsOut { double* o; int64_t n; int64_t p; struct Ref { double* a; NPS_HD void operator=(double v) { *a = v; } };
                        __device__ Ref operator[](int i) { return Ref{o + (int64_t)i * n + p}; } } out{obs, n, p};
        plant_observe(st, prm, out);
    }
    if (reward) reward[p] = plant_reward(st, prm);
}

// ------------------------------------------------------------------------------------------------
// threshold flag kernel: one thread per (plant, flag word) evaluates the live threshold rows of that word; cooldown stamps in SoA;
// __ballot_sync compacts "this plant fired something" into one word per warp for the host drain.
// ------------------------------------------------------------------------------------------------
// Logged columns that are pure functions of carried fields.  code = -(2 + 4 * kind + unit):
//   kind 0  sum_wear_level of feedwater pump `unit` = impeller + max(motor, pump, thrust bearing) + seal wear
//           (FeedwaterPumpLubricationSystem.get_state_dict: feedwater/pump_lubrication.py:1585-1596)
__device__ __forceinline__ double threshold_derived(const double* __restrict__ slab, int64_t n, int64_t p, int code) {
    const int k = -code - 2, kind = k >> 2, unit = k & 3;
    if (kind == 0) {
        const int f0 = (int)((offsetof(PlantState, fw) + offsetof(FeedwaterState, pump) + unit * sizeof(FWPumpState) +
                              offsetof(FWPumpState, lub) + offsetof(LubCore, component_wear)) / sizeof(double));
        const double* w = slab + (int64_t)f0 * n + p;
        const double imp = w[0], mb = w[(int64_t)FWL_MOTOR_BRG * n], pb = w[(int64_t)FWL_PUMP_BRG * n];
        const double tb = w[(int64_t)FWL_THRUST_BRG * n], sw = w[(int64_t)FWL_SEALS * n];
        return imp + py_max3(mb, pb, tb) + sw;
    }
    return NAN;
}


constexpr int kThrRowsPerThread = 8;
__global__ void __launch_bounds__(128)
nps_threshold_kernel(const double* __restrict__ slab, const Threshold* __restrict__ live, const int32_t* __restrict__ word_start,
                     int time_field, double* __restrict__ last_fired, uint32_t* __restrict__ flags,
                     uint32_t* __restrict__ any_warp, int64_t n) {
    // thread = (plant, flag word).  A flag word covers 32 consecutive rows of the caller's table; the thread evaluates
    // the LIVE rows among them (inert rows never reach the device) and OWNS that word: one plain coalesced store, no
    // atomics and no clearing pass.  The rows of the word are staged in shared memory once per block; the cooldown
    // stamps and values of 8 rows at a time are independent loads issued back to back, so the kernel streams instead of
    // chasing 2 x 90 dependent loads per plant (that first version ran at 11 % of the HBM roofline; profiles/).
    __shared__ Threshold rows[32];
    const int w = blockIdx.y;
    const int ws = word_start[w], cnt = word_start[w + 1] - ws;
    if ((int)threadIdx.x < cnt) rows[threadIdx.x] = live[ws + threadIdx.x];
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bits = 0;
    if (p < n && cnt > 0) {
        const double now = slab[(int64_t)time_field * n + p];
        for (int j0 = 0; j0 < cnt; j0 += kThrRowsPerThread) {
            double last[kThrRowsPerThread], val[kThrRowsPerThread];
#pragma unroll
            for (int j = 0; j < kThrRowsPerThread; ++j) {
                if (j0 + j < cnt) {
                    const Threshold& th = rows[j0 + j];
                    last[j] = last_fired[(int64_t)th.row * n + p];
                    val[j] = (th.field >= 0) ? slab[(int64_t)th.field * n + p] : threshold_derived(slab, n, p, th.field);
                }
            }
#pragma unroll
            for (int j = 0; j < kThrRowsPerThread; ++j) {
                if (j0 + j >= cnt) break;
                const Threshold& th = rows[j0 + j];
                // _is_threshold_in_cooldown: state_manager.py:1267-1305 (checked before the value is looked at)
                if ((now - last[j]) < th.cooldown) continue;
                const double v = val[j];
                bool fire;
                switch (th.cmp) {   // _check_threshold_condition: state_manager.py:1412-1442
                    case 0: fire = v > th.value; break;
                    case 1: fire = v < th.value; break;
                    case 2: fire = v >= th.value; break;
                    case 3: fire = v <= th.value; break;
                    case 4: fire = fabs(v - th.value) < 1e-3; break;
                    case 5: fire = fabs(v - th.value) >= 1e-3; break;
                    default: fire = false; break;
                }
                if (fire) {
                    last_fired[(int64_t)th.row * n + p] = now;   // _record_threshold_violation_time
                    bits |= 1u << (th.row & 31);
                }
            }
        }
    }
    if (p < n) flags[(int64_t)w * n + p] = bits;
    const unsigned ballot = __ballot_sync(0xffffffffu, bits != 0);
    if ((threadIdx.x & 31) == 0 && ballot && any_warp) atomicOr(&any_warp[p >> 5], ballot);
}

// Event-list form of the same check: thread per plant walks the live rows (staged in shared memory; every row's value
// is one coalesced load across the warp) and appends violations to the event list exactly as the in-launch monitor does.
__global__ void __launch_bounds__(128)
nps_threshold_events_kernel(const double* __restrict__ slab, const Threshold* __restrict__ live, int n_live, int time_field,
                            double* __restrict__ last_fired, nps_event* __restrict__ events, uint32_t* __restrict__ n_events,
                            uint32_t event_cap, int32_t step, int64_t n) {
    extern __shared__ Threshold s_live[];
    for (int j = threadIdx.x; j < n_live; j += blockDim.x) s_live[j] = live[j];
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n;
    const int64_t pc = valid ? p : n - 1;             // out-of-range lanes read the last plant and never fire
    const unsigned lane = threadIdx.x & 31u;
    const double now = slab[(int64_t)time_field * n + pc];
    // eight rows at a time: their values are independent coalesced loads issued together, the warp votes once per
    // group, and only a group with a violation walks its rows to stamp cooldowns and append events
    constexpr int kChunk = 8;
    for (int j0 = 0; j0 < n_live; j0 += kChunk) {
        double v[kChunk];
        unsigned hit = 0;
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const int f = (j0 + j < n_live) ? s_live[j0 + j].field : 0;
            v[j] = slab[(int64_t)(f >= 0 ? f : 0) * n + pc];
        }
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            if (j0 + j >= n_live) break;
            const Threshold& th = s_live[j0 + j];
            if (th.field < 0) v[j] = threshold_derived(slab, n, pc, th.field);
            if (valid && threshold_compare(th.cmp, v[j], th.value)) hit |= 1u << j;
        }
        if (!__ballot_sync(0xffffffffu, hit != 0)) continue;
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            if (j0 + j >= n_live) break;
            const Threshold th = s_live[j0 + j];
            bool fire = (hit >> j) & 1u;
            if (fire) {
                double* stamp = last_fired + (int64_t)th.row * n + p;
                if ((now - *stamp) < th.cooldown) fire = false;
                else *stamp = now;
            }
            const unsigned m = __ballot_sync(0xffffffffu, fire);
            if (m) {
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(n_events, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (fire) {
                    const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (slot < event_cap) events[slot] = nps_event{(int32_t)p, th.row, step, 0, v[j], now};
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// trajectory ring-buffer row: gather the logged fields of a tile of plants through shared memory
// ------------------------------------------------------------------------------------------------
constexpr int kLogTilePlants = 128;
constexpr int kLogTileFields = 32;
__global__ void nps_log_row_kernel(const double* __restrict__ slab, const int32_t* __restrict__ fields, int n_logged,
                                   double* __restrict__ ring_row, int64_t n) {
    __shared__ double tile[kLogTileFields][kLogTilePlants + 1];
    const int64_t p0 = (int64_t)blockIdx.x * kLogTilePlants;
    const int f0 = blockIdx.y * kLogTileFields;
    const int tx = threadIdx.x;   // plant within tile
    for (int j = 0; j < kLogTileFields; ++j) {
        const int lf = f0 + j;
        if (lf < n_logged && p0 + tx < n) tile[j][tx] = slab[(int64_t)fields[lf] * n + p0 + tx];
    }
    __syncthreads();
    for (int j = 0; j < kLogTileFields; ++j) {
        const int lf = f0 + j;
        if (lf < n_logged && p0 + tx < n) ring_row[(int64_t)lf * n + p0 + tx] = tile[j][tx];
    }
}

// ------------------------------------------------------------------------------------------------
// trajectory ring-buffer row, TMA path (sm_100a): a logged field of all plants is one contiguous row of the slab, so a
// ring row is n_logged row copies.  Each CTA is one warp whose lane 0 drives a kLogStages-deep pipeline of
// cp.async.bulk transfers: global -> shared (completion on an mbarrier) and shared -> global (bulk async-group); no
// thread touches the data.  Work items are (field, 16 KB chunk of plants), dealt round-robin to a persistent grid of
// kLogCtasPerSm CTAs per SM.  Needs 16-byte aligned rows, i.e. an even plant count; otherwise nps_log_row_kernel runs.
// ------------------------------------------------------------------------------------------------
constexpr int kLogStageBytes = 16384;
constexpr int kLogStages = 4;
constexpr int kLogCtasPerSm = 3;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32) nps_log_row_tma_kernel(const double* __restrict__ slab, const int32_t* __restrict__ fields,
                                                             int n_logged, double* __restrict__ ring_row, int64_t n,
                                                             int chunks_per_field) {
    extern __shared__ __align__(128) unsigned char stage_mem[];
    __shared__ uint64_t bar[kLogStages];
    if (threadIdx.x != 0) return;                      // one driver thread; the copies are done by the TMA unit
    for (int s = 0; s < kLogStages; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int64_t total = (int64_t)n_logged * chunks_per_field;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t count = (total > first) ? (total - first + stride - 1) / stride : 0;
    constexpr int64_t kChunkPlants = kLogStageBytes / 8;
    auto locate = [&](int64_t k, const double*& src, double*& dst, uint32_t& bytes) {
        const int64_t item = first + k * stride;
        const int f = (int)(item / chunks_per_field);
        const int64_t p0 = (item - (int64_t)f * chunks_per_field) * kChunkPlants;
        const int64_t np = (n - p0 < kChunkPlants) ? (n - p0) : kChunkPlants;
        src = slab + (int64_t)fields[f] * n + p0;
        dst = ring_row + (int64_t)f * n + p0;
        bytes = (uint32_t)(np * 8);
    };
    auto load = [&](int64_t k) {
        const double* src; double* dst; uint32_t bytes;
        locate(k, src, dst, bytes);
        const int s = (int)(k % kLogStages);
        mbar_expect_tx(&bar[s], bytes);
        bulk_g2s(stage_mem + (size_t)s * kLogStageBytes, src, bytes, &bar[s]);
    };
    for (int64_t k = 0; k < count && k < kLogStages; ++k) load(k);
    for (int64_t k = 0; k < count; ++k) {
        const int s = (int)(k % kLogStages);
        mbar_wait(&bar[s], (uint32_t)((k / kLogStages) & 1));
        const double* src; double* dst; uint32_t bytes;
        locate(k, src, dst, bytes);
        bulk_s2g(dst, stage_mem + (size_t)s * kLogStageBytes, bytes);
        // refill the stage whose store was issued one iteration ago (its shared-memory read has had time to finish)
        if (k >= 1 && k - 1 + kLogStages < count) {
            bulk_wait_read<1>();
            load(k - 1 + kLogStages);
        }
    }
    bulk_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the kernel ends
}

// ------------------------------------------------------------------------------------------------
// maintenance effects: one thread per affected plant applies that plant's requests in order.  Sparse by nature
// (work orders fire for a handful of plants per check), so the whole record is loaded and stored.
// ------------------------------------------------------------------------------------------------
struct MaintRequest { int32_t plant, target, action, arg; };
__global__ void nps_maintenance_kernel(double* __restrict__ slab, const __grid_constant__ PlantParams prm,
                                       const MaintRequest* __restrict__ req, const int32_t* __restrict__ group_start,
                                       int n_groups, int32_t* __restrict__ status, int64_t n) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int lo = group_start[g], hi = group_start[g + 1];
    const int64_t p = req[lo].plant;
    PlantState st;
    double* sv = reinterpret_cast<double*>(&st);
    for (int f = 0; f < kNState; ++f) sv[f] = slab[(int64_t)f * n + p];
    for (int i = lo; i < hi; ++i) status[i] = maintenance_apply(st, prm, req[i].target, req[i].action, req[i].arg);
    for (int f = 0; f < kNState; ++f) slab[(int64_t)f * n + p] = sv[f];
}

__global__ void nps_gather_kernel(const double* __restrict__ slab, const int32_t* __restrict__ fields, int n_fields,
                                  double* __restrict__ out, int64_t n) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    for (int j = 0; j < n_fields; ++j) out[(int64_t)j * n + p] = slab[(int64_t)fields[j] * n + p];
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int nps_abi_version(void) { return NPS_ABI_VERSION; }
const char* nps_last_error(void) { return g_last_error.c_str(); }
int nps_n_state(void) { return kNState; }
int nps_n_params(void) { return kNParams; }
const char* nps_field_name(int f) { return (f >= 0 && f < kNState) ? kStateFieldNames[f] : nullptr; }
const char* nps_param_name(int f) { return (f >= 0 && f < kNParams) ? kParamFieldNames[f] : nullptr; }

int nps_create(int64_t n_plants, int device, nps_handle** out) {
    if (!out || n_plants <= 0) return fail("nps_create: bad arguments");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return fail("nps_create: no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= count) return fail("nps_create: device index out of range");
    DeviceGuard guard(device);
    nps_handle* h = new nps_handle();
    h->n = n_plants; h->device = device;
    std::memset(&h->params, 0, sizeof(PlantParams));
    cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device);
    { const char* e = getenv("NPS_LOG_ROW_TILE"); h->log_row_tile_only = e && e[0] == '1'; }
    { const char* e = getenv("NPS_SMALL_SHAPE"); h->small_shape = (e && e[0] == '1') ? 1 : 0; }
    { const char* e = getenv("NPS_LARGE_BATCH"); if (e && atoll(e) > 0) h->large_batch = atoll(e); }
    NPS_CUDA(cudaFuncSetAttribute(nps_log_row_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLogStages * kLogStageBytes));
    // the step kernel keeps one PlantState per thread in local memory
#if defined(NPS_STEP_BLOCK)
    NPS_CUDA(cudaFuncSetCacheConfig(nps_step_kernel<NPS_STEP_BLOCK, NPS_STEP_MINBLOCKS>, cudaFuncCachePreferL1));
#else
    NPS_CUDA(cudaFuncSetCacheConfig(nps_step_kernel<448, 1>, cudaFuncCachePreferL1));
    // one-warp blocks: several of them share an SM, each with its 4 KB threshold-row stage; PreferL1 alone would leave
    // shared memory for a single block per SM (measured: 8 192 plants ran as two waves)
    NPS_CUDA(cudaFuncSetAttribute(nps_step_kernel<32, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 25));
    NPS_CUDA(cudaFuncSetAttribute(nps_step_split_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 40));
#endif
    *out = h;
    return 0;
}

void nps_destroy(nps_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaFree(h->d_setpoint); cudaFree(h->d_action); cudaFree(h->d_mag); cudaFree(h->d_noise);
    cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_done); cudaFree(h->d_thresholds); cudaFree(h->d_word_start);
    cudaFree(h->d_logged); cudaFree(h->d_gather_fields); cudaFree(h->d_gather_out);
    cudaFree(h->mb.d_req); cudaFree(h->mb.d_start); cudaFree(h->mb.d_status);
    cudaFreeHost(h->mb.h_req); cudaFreeHost(h->mb.h_start); cudaFreeHost(h->mb.h_status);
    for (auto& q : h->pipe) {
        cudaFree(q.d_action); cudaFree(q.d_mag); cudaFree(q.d_noise); cudaFree(q.d_setpoint);
        cudaFree(q.d_obs); cudaFree(q.d_reward); cudaFree(q.d_done);
        if (q.in_done) cudaEventDestroy(q.in_done);
        if (q.kernel_done) cudaEventDestroy(q.kernel_done);
        if (q.out_done) cudaEventDestroy(q.out_done);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    delete h;
}

int64_t nps_n_plants(const nps_handle* h) { return h ? h->n : 0; }

int nps_set_params(nps_handle* h, const double* params_host, int n_params) {
    if (!h || !params_host) return fail("nps_set_params: null argument");
    if (n_params != kNParams) return fail("nps_set_params: parameter count mismatch");
    std::memcpy(&h->params, params_host, sizeof(PlantParams));   // passed to kernels by value (constant bank)
    return 0;
}

static int launch_step(nps_handle* h, const StepArgs& a, cudaStream_t s) {
#if defined(NPS_STEP_BLOCK)
    nps_step_kernel<NPS_STEP_BLOCK, NPS_STEP_MINBLOCKS><<<(int)((h->n + NPS_STEP_BLOCK - 1) / NPS_STEP_BLOCK), NPS_STEP_BLOCK, 0, s>>>(h->params, a);
#else
    if (h->n >= h->large_batch) nps_step_kernel<448, 1><<<(int)((h->n + 447) / 448), 448, 0, s>>>(h->params, a);
    else {
        const int blocks = (int)((h->n + kSplitPlants - 1) / kSplitPlants);
        if (h->small_shape == 0 && blocks <= h->n_sms * 4) nps_step_split_kernel<<<blocks, 2 * kSplitPlants, 0, s>>>(h->params, a);
        else nps_step_kernel<32, 1><<<(int)((h->n + 31) / 32), 32, 0, s>>>(h->params, a);
    }
#endif
    NPS_CUDA(cudaGetLastError());
    return 0;
}

int nps_step_monitored(nps_handle* h, double* d_state, const int8_t* d_action, const double* d_magnitude, const double* d_noise,
                       const double* d_setpoint, int k_substeps, double* d_obs, double* d_reward, uint8_t* d_done,
                       const nps_monitor* mon, void* cuda_stream) {
    if (!h || !d_state) return fail("nps_step: null argument");
    if (k_substeps <= 0) return fail("nps_step: k_substeps must be positive");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    StepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.slab = d_state; a.action = d_action; a.magnitude = d_magnitude; a.noise = d_noise; a.setpoint = d_setpoint;
    a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.n = h->n; a.k_substeps = k_substeps;
    a.rng = h->rng;
    if (a.rng.enabled && !d_noise) h->rng.step0 += (uint64_t)k_substeps;   // the next launch continues the stream
    if (mon) {
        MonitorArgs& m = a.mon;
        if (mon->d_last_fired) {
            if (h->n_thresholds == 0) return fail("nps_step_monitored: d_last_fired given but no thresholds set");
            if (!mon->d_events || !mon->d_n_events || mon->event_capacity == 0) return fail("nps_step_monitored: threshold evaluation needs an event list");
        }
        if (mon->n_watch < 0 || mon->n_watch > 32) return fail("nps_step_monitored: at most 32 watched fields");
        if (mon->n_watch > 0 && (!mon->d_watch_fields || !mon->d_watch_step)) return fail("nps_step_monitored: null watch arrays");
        m.enabled = 1;
        m.n_live = mon->d_last_fired ? h->n_live_thresholds : 0;
        m.skip_last_check = mon->skip_last_check;
        m.n_watch = mon->n_watch;
        m.live = h->d_thresholds; m.last_fired = mon->d_last_fired;
        m.events = mon->d_events; m.n_events = mon->d_n_events; m.event_cap = mon->event_capacity;
        m.watch_fields = mon->d_watch_fields; m.watch_step = mon->d_watch_step;
        m.first_scram_step = mon->d_first_scram_step; m.first_nan_reset_step = mon->d_first_nan_reset_step; m.status = mon->d_status;
        m.reward_k = mon->d_reward_k; m.done_k = mon->d_done_k; m.step0 = mon->step0;
    }
    return launch_step(h, a, s);
}

int nps_step(nps_handle* h, double* d_state, const int8_t* d_action, const double* d_magnitude, const double* d_noise,
             const double* d_setpoint, int k_substeps, double* d_obs, double* d_reward, uint8_t* d_done,
             void* cuda_stream) {
    return nps_step_monitored(h, d_state, d_action, d_magnitude, d_noise, d_setpoint, k_substeps, d_obs, d_reward, d_done,
                              nullptr, cuda_stream);
}

static int ensure_staging(nps_handle* h, int k) {
    if (h->staged_k >= k && h->d_obs) return 0;
    cudaFree(h->d_action); cudaFree(h->d_mag); cudaFree(h->d_noise); cudaFree(h->d_setpoint);
    h->d_action = nullptr; h->d_mag = nullptr; h->d_noise = nullptr; h->d_setpoint = nullptr;
    NPS_CUDA(cudaMalloc(&h->d_action, (size_t)k * h->n));
    NPS_CUDA(cudaMalloc(&h->d_mag, (size_t)k * h->n * sizeof(double)));
    NPS_CUDA(cudaMalloc(&h->d_noise, (size_t)k * NPS_NOISE_PER_STEP * h->n * sizeof(double)));
    NPS_CUDA(cudaMalloc(&h->d_setpoint, (size_t)k * h->n * sizeof(double)));
    if (!h->d_obs) {
        NPS_CUDA(cudaMalloc(&h->d_obs, (size_t)NPS_OBS_DIM * h->n * sizeof(double)));
        NPS_CUDA(cudaMalloc(&h->d_reward, (size_t)h->n * sizeof(double)));
        NPS_CUDA(cudaMalloc(&h->d_done, (size_t)h->n));
    }
    h->staged_k = k;
    return 0;
}

int nps_step_host(nps_handle* h, double* d_state, const int8_t* h_action, const double* h_magnitude,
                  const double* h_noise, const double* h_setpoint, int k_substeps, double* h_obs, double* h_reward,
                  uint8_t* h_done, void* cuda_stream) {
    if (!h || !d_state) return fail("nps_step_host: null argument");
    if (k_substeps <= 0) return fail("nps_step_host: k_substeps must be positive");
    DeviceGuard guard(h->device);
    if (ensure_staging(h, k_substeps)) return -1;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t kn = (size_t)k_substeps * h->n;
    if (h_action) NPS_CUDA(cudaMemcpyAsync(h->d_action, h_action, kn, cudaMemcpyHostToDevice, s));
    if (h_magnitude) NPS_CUDA(cudaMemcpyAsync(h->d_mag, h_magnitude, kn * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_noise) NPS_CUDA(cudaMemcpyAsync(h->d_noise, h_noise, kn * NPS_NOISE_PER_STEP * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_setpoint) NPS_CUDA(cudaMemcpyAsync(h->d_setpoint, h_setpoint, kn * sizeof(double), cudaMemcpyHostToDevice, s));
    if (nps_step(h, d_state, h_action ? h->d_action : nullptr, h_magnitude ? h->d_mag : nullptr,
                 h_noise ? h->d_noise : nullptr, h_setpoint ? h->d_setpoint : nullptr, k_substeps, h_obs ? h->d_obs : nullptr,
                 h_reward ? h->d_reward : nullptr, h_done ? h->d_done : nullptr, cuda_stream)) return -1;
    if (h_obs) NPS_CUDA(cudaMemcpyAsync(h_obs, h->d_obs, (size_t)NPS_OBS_DIM * h->n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_reward) NPS_CUDA(cudaMemcpyAsync(h_reward, h->d_reward, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_done) NPS_CUDA(cudaMemcpyAsync(h_done, h->d_done, (size_t)h->n, cudaMemcpyDeviceToHost, s));
    NPS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

static int ensure_pipe(nps_handle* h, int k) {
    if (h->pipe_k >= k && h->copy_stream) return 0;
    if (!h->copy_stream) NPS_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->out_stream) NPS_CUDA(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    for (auto& q : h->pipe) {
        cudaFree(q.d_action); cudaFree(q.d_mag); cudaFree(q.d_noise); cudaFree(q.d_setpoint);
        NPS_CUDA(cudaMalloc(&q.d_action, (size_t)k * h->n));
        NPS_CUDA(cudaMalloc(&q.d_mag, (size_t)k * h->n * sizeof(double)));
        NPS_CUDA(cudaMalloc(&q.d_noise, (size_t)k * NPS_NOISE_PER_STEP * h->n * sizeof(double)));
        NPS_CUDA(cudaMalloc(&q.d_setpoint, (size_t)k * h->n * sizeof(double)));
        if (!q.d_obs) {
            NPS_CUDA(cudaMalloc(&q.d_obs, (size_t)NPS_OBS_DIM * h->n * sizeof(double)));
            NPS_CUDA(cudaMalloc(&q.d_reward, (size_t)h->n * sizeof(double)));
            NPS_CUDA(cudaMalloc(&q.d_done, (size_t)h->n));
            const char* tr = getenv("NPS_PIPE_TRACE");
            h->pipe_trace = tr && tr[0] == '1';
            const unsigned fl = h->pipe_trace ? cudaEventDefault : cudaEventDisableTiming;
            NPS_CUDA(cudaEventCreateWithFlags(&q.in_done, fl));
            NPS_CUDA(cudaEventCreateWithFlags(&q.kernel_done, fl));
            NPS_CUDA(cudaEventCreateWithFlags(&q.out_done, fl));
            if (h->pipe_trace) { NPS_CUDA(cudaEventCreate(&q.t_in0)); NPS_CUDA(cudaEventCreate(&q.t_k0)); }
        }
    }
    h->pipe_k = k;
    return 0;
}

int nps_step_host_async(nps_handle* h, double* d_state, const int8_t* h_action, const double* h_magnitude,
                        const double* h_noise, const double* h_setpoint, int k_substeps, double* h_obs, double* h_reward,
                        uint8_t* h_done, void* cuda_stream) {
    if (!h || !d_state) return fail("nps_step_host_async: null argument");
    if (k_substeps <= 0) return fail("nps_step_host_async: k_substeps must be positive");
    DeviceGuard guard(h->device);
    if (ensure_pipe(h, k_substeps)) return -1;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const int slot = (int)(h->pipe_count++ % NPS_PIPE_DEPTH);
    nps_handle::Pipe& q = h->pipe[slot];
    const size_t kn = (size_t)k_substeps * h->n;
    // the staging set may still be read by the launch issued NPS_PIPE_DEPTH calls ago
    NPS_CUDA(cudaStreamWaitEvent(h->copy_stream, q.kernel_done, 0));
    if (h->pipe_trace) NPS_CUDA(cudaEventRecord(q.t_in0, h->copy_stream));
    if (h_action) NPS_CUDA(cudaMemcpyAsync(q.d_action, h_action, kn, cudaMemcpyHostToDevice, h->copy_stream));
    if (h_magnitude) NPS_CUDA(cudaMemcpyAsync(q.d_mag, h_magnitude, kn * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    if (h_noise) NPS_CUDA(cudaMemcpyAsync(q.d_noise, h_noise, kn * NPS_NOISE_PER_STEP * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    if (h_setpoint) NPS_CUDA(cudaMemcpyAsync(q.d_setpoint, h_setpoint, kn * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    NPS_CUDA(cudaEventRecord(q.in_done, h->copy_stream));
    NPS_CUDA(cudaStreamWaitEvent(s, q.in_done, 0));
    if (h->pipe_trace) NPS_CUDA(cudaEventRecord(q.t_k0, s));
    if (nps_step(h, d_state, h_action ? q.d_action : nullptr, h_magnitude ? q.d_mag : nullptr, h_noise ? q.d_noise : nullptr,
                 h_setpoint ? q.d_setpoint : nullptr, k_substeps, h_obs ? q.d_obs : nullptr, h_reward ? q.d_reward : nullptr,
                 h_done ? q.d_done : nullptr, cuda_stream)) return -1;
    NPS_CUDA(cudaEventRecord(q.kernel_done, s));
    // results go home on their own stream: neither the next launch nor the next input copy queues behind them
    NPS_CUDA(cudaStreamWaitEvent(h->out_stream, q.kernel_done, 0));
    if (h_obs) NPS_CUDA(cudaMemcpyAsync(h_obs, q.d_obs, (size_t)NPS_OBS_DIM * h->n * sizeof(double), cudaMemcpyDeviceToHost, h->out_stream));
    if (h_reward) NPS_CUDA(cudaMemcpyAsync(h_reward, q.d_reward, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, h->out_stream));
    if (h_done) NPS_CUDA(cudaMemcpyAsync(h_done, q.d_done, (size_t)h->n, cudaMemcpyDeviceToHost, h->out_stream));
    NPS_CUDA(cudaEventRecord(q.out_done, h->out_stream));
    return slot;
}

__global__ void nps_selftest_pow_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = py_pow(x[i], y[i]);
}

int nps_selftest_pow(const double* d_x, const double* d_y, double* d_out, int64_t n, void* cuda_stream) {
    if (!d_x || !d_y || !d_out || n < 0) return fail("nps_selftest_pow: bad arguments");
    if (n == 0) return 0;
    nps_selftest_pow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_x, d_y, d_out, n);
    NPS_CUDA(cudaGetLastError());
    return 0;
}

// FP64 peak microbenchmark (SURVEY 6 / BASELINE.md 2): kChains independent DFMA chains per thread, full occupancy.
// -fmad=false does not touch explicit __fma_rn, so every loop iteration is kChains DFMA instructions.
constexpr int kFmaChains = 8;
__global__ void __launch_bounds__(256) nps_fp64_fma_kernel(double* __restrict__ out, int iters, double a, double b) {
    double x[kFmaChains];
#pragma unroll
    for (int c = 0; c < kFmaChains; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kFmaChains; ++c) x[c] = __fma_rn(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kFmaChains; ++c) s += x[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive, never true in practice
}

int nps_measure_fp64_peak(int device, int iters, double* out_tflops, double* out_ms) {
    if (!out_tflops || iters <= 0) return fail("nps_measure_fp64_peak: bad arguments");
    DeviceGuard guard(device);
    int sms = 0;
    NPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256;      // 2 048 threads per SM: every scheduler has 16 warps of DFMA work
    double* d_out = nullptr;
    NPS_CUDA(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    NPS_CUDA(cudaEventCreate(&e0)); NPS_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {   // rep 0 is the warm-up
        NPS_CUDA(cudaEventRecord(e0, 0));
        nps_fp64_fma_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
        NPS_CUDA(cudaEventRecord(e1, 0));
        NPS_CUDA(cudaEventSynchronize(e1));
        float ms = 0; NPS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    NPS_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    const double flops = 2.0 * kFmaChains * (double)iters * blocks * threads;
    *out_tflops = flops / (best * 1e-3) / 1e12;
    if (out_ms) *out_ms = best;
    return 0;
}

int nps_set_small_batch_shape(nps_handle* h, int shape) {
    if (!h || (shape != 0 && shape != 1)) return fail("nps_set_small_batch_shape: shape is 0 (two threads per plant) or 1 (one thread per plant)");
    h->small_shape = shape;
    return 0;
}

int nps_pipe_depth(void) { return NPS_PIPE_DEPTH; }

int nps_set_device_rng(nps_handle* h, int enabled, uint64_t seed, uint64_t plant_offset, uint64_t first_step) {
    if (!h) return fail("nps_set_device_rng: null handle");
    h->rng.enabled = enabled ? 1 : 0;
    h->rng.seed = seed; h->rng.plant_offset = plant_offset; h->rng.step0 = first_step;
    return 0;
}

int nps_device_rng_draws(uint64_t seed, uint64_t plant, uint64_t step, double* out5) {
    if (!out5) return fail("nps_device_rng_draws: null output");
    const StepDraws d = plant_step_draws(seed, plant, step);   // host evaluation of the same generator (libm transcendentals)
    out5[0] = d.z_heat; out5[1] = d.z_ph; out5[2] = d.u_ph[0]; out5[3] = d.u_ph[1]; out5[4] = d.u_ph[2];
    return 0;
}

int nps_wait(nps_handle* h, int ticket) {
    if (!h || ticket < 0 || ticket >= NPS_PIPE_DEPTH || !h->pipe[ticket].out_done) return fail("nps_wait: bad ticket");
    DeviceGuard guard(h->device);
    NPS_CUDA(cudaEventSynchronize(h->pipe[ticket].out_done));
    if (h->pipe_trace) {   // where one pipelined step spent its time on the device
        nps_handle::Pipe& q = h->pipe[ticket];
        float copy_ms = 0, gap_ms = 0, kern_ms = 0, out_ms = 0;
        cudaEventElapsedTime(&copy_ms, q.t_in0, q.in_done);
        cudaEventElapsedTime(&gap_ms, q.in_done, q.t_k0);
        cudaEventElapsedTime(&kern_ms, q.t_k0, q.kernel_done);
        cudaEventElapsedTime(&out_ms, q.kernel_done, q.out_done);
        fprintf(stderr, "[nps pipe] slot %d in %.2f gap %.2f kernel %.2f out %.2f ms\n", ticket, copy_ms, gap_ms, kern_ms, out_ms);
    }
    return 0;
}

int nps_observe(nps_handle* h, const double* d_state, double* d_obs, double* d_reward, void* cuda_stream) {
    if (!h || !d_state) return fail("nps_observe: null argument");
    DeviceGuard guard(h->device);
    const int block = 64;
    const int grid = (int)((h->n + block - 1) / block);
    nps_observe_kernel<<<grid, block, 0, (cudaStream_t)cuda_stream>>>(d_state, h->params, h->n, d_obs, d_reward);
    NPS_CUDA(cudaGetLastError());
    return 0;
}

int nps_set_thresholds(nps_handle* h, const int32_t* field, const int32_t* comparator, const double* value,
                       const double* cooldown_minutes, int n_thresholds) {
    if (!h || n_thresholds < 0) return fail("nps_set_thresholds: bad arguments");
    DeviceGuard guard(h->device);
    cudaFree(h->d_thresholds); h->d_thresholds = nullptr; h->n_thresholds = 0;
    cudaFree(h->d_word_start); h->d_word_start = nullptr;
    if (n_thresholds == 0) return 0;
    std::vector<Threshold> t;      // live rows only (field != -1), each remembering its row index in the caller's table
    for (int i = 0; i < n_thresholds; ++i) {
        if (field[i] >= kNState) return fail("nps_set_thresholds: field index out of range");
        if (field[i] != -1) t.push_back(Threshold{field[i], comparator[i], value[i], cooldown_minutes[i], i, 0});
    }
    if (!t.empty()) {
        NPS_CUDA(cudaMalloc(&h->d_thresholds, sizeof(Threshold) * t.size()));
        NPS_CUDA(cudaMemcpy(h->d_thresholds, t.data(), sizeof(Threshold) * t.size(), cudaMemcpyHostToDevice));
        const int n_words = (n_thresholds + 31) / 32;
        std::vector<int32_t> ws(n_words + 1, 0);
        for (const Threshold& th : t) ws[th.row / 32 + 1]++;
        for (int w = 0; w < n_words; ++w) ws[w + 1] += ws[w];
        NPS_CUDA(cudaMalloc(&h->d_word_start, sizeof(int32_t) * ws.size()));
        NPS_CUDA(cudaMemcpy(h->d_word_start, ws.data(), sizeof(int32_t) * ws.size(), cudaMemcpyHostToDevice));
    }
    h->n_thresholds = n_thresholds;
    h->n_live_thresholds = (int)t.size();
    return 0;
}

int nps_check_thresholds(nps_handle* h, const double* d_state, double* d_last_fired, uint32_t* d_flags,
                         uint32_t* d_any_warp, void* cuda_stream) {
    if (!h || !d_state || !d_last_fired || !d_flags) return fail("nps_check_thresholds: null argument");
    if (h->n_thresholds == 0) return fail("nps_check_thresholds: no thresholds set");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const int n_words = (h->n_thresholds + 31) / 32;
    if (d_any_warp) NPS_CUDA(cudaMemsetAsync(d_any_warp, 0, sizeof(uint32_t) * (size_t)((h->n + 31) / 32), s));
    if (h->n_live_thresholds == 0) {   // nothing can fire: all flag words are zero
        NPS_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * (size_t)n_words * h->n, s));
        return 0;
    }
    const int block = 128;
    dim3 grid((unsigned)((h->n + block - 1) / block), (unsigned)n_words);   // every flag word is written by its owner
    nps_threshold_kernel<<<grid, block, 0, s>>>(d_state, h->d_thresholds, h->d_word_start, kTimeMinutesField, d_last_fired,
                                               d_flags, d_any_warp, h->n);
    NPS_CUDA(cudaGetLastError());
    return 0;
}

int nps_check_thresholds_events(nps_handle* h, const double* d_state, double* d_last_fired, nps_event* d_events,
                                uint32_t* d_n_events, uint32_t event_capacity, int32_t step, void* cuda_stream) {
    if (!h || !d_state || !d_last_fired || !d_events || !d_n_events || event_capacity == 0) return fail("nps_check_thresholds_events: null argument");
    if (h->n_thresholds == 0) return fail("nps_check_thresholds_events: no thresholds set");
    if (h->n_live_thresholds == 0) return 0;
    DeviceGuard guard(h->device);
    const int block = 128;
    const size_t smem = sizeof(Threshold) * (size_t)h->n_live_thresholds;
    if (smem > 48 * 1024) return fail("nps_check_thresholds_events: too many live threshold rows for one shared-memory stage");
    nps_threshold_events_kernel<<<(unsigned)((h->n + block - 1) / block), block, smem, (cudaStream_t)cuda_stream>>>(
        d_state, h->d_thresholds, h->n_live_thresholds, kTimeMinutesField, d_last_fired, d_events, d_n_events, event_capacity,
        step, h->n);
    NPS_CUDA(cudaGetLastError());
    return 0;
}

int nps_set_logged_fields(nps_handle* h, const int32_t* fields, int n_logged) {
    if (!h || n_logged < 0) return fail("nps_set_logged_fields: bad arguments");
    DeviceGuard guard(h->device);
    cudaFree(h->d_logged); h->d_logged = nullptr; h->n_logged = 0;
    if (n_logged == 0) return 0;
    for (int i = 0; i < n_logged; ++i) if (fields[i] < 0 || fields[i] >= kNState) return fail("nps_set_logged_fields: field index out of range");
    NPS_CUDA(cudaMalloc(&h->d_logged, sizeof(int32_t) * n_logged));
    NPS_CUDA(cudaMemcpy(h->d_logged, fields, sizeof(int32_t) * n_logged, cudaMemcpyHostToDevice));
    h->n_logged = n_logged;
    return 0;
}

int nps_log_row(nps_handle* h, const double* d_state, double* d_ring, int64_t ring_rows, int64_t write_index,
                void* cuda_stream) {
    if (!h || !d_state || !d_ring || ring_rows <= 0) return fail("nps_log_row: bad arguments");
    if (h->n_logged == 0) return fail("nps_log_row: no logged fields set");
    DeviceGuard guard(h->device);
    double* row = d_ring + (write_index % ring_rows) * (int64_t)h->n_logged * h->n;
    const bool aligned = (h->n % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_state) | reinterpret_cast<uintptr_t>(d_ring)) % 16 == 0);
    if (aligned && !h->log_row_tile_only) {
        const int chunks = (int)((h->n * 8 + kLogStageBytes - 1) / kLogStageBytes);
        const int64_t items = (int64_t)h->n_logged * chunks;
        int grid = h->n_sms * kLogCtasPerSm;
        if (grid > items) grid = (int)items;
        nps_log_row_tma_kernel<<<grid, 32, kLogStages * kLogStageBytes, (cudaStream_t)cuda_stream>>>(
            d_state, h->d_logged, h->n_logged, row, h->n, chunks);
    } else {
        dim3 grid((unsigned)((h->n + kLogTilePlants - 1) / kLogTilePlants), (unsigned)((h->n_logged + kLogTileFields - 1) / kLogTileFields));
        nps_log_row_kernel<<<grid, kLogTilePlants, 0, (cudaStream_t)cuda_stream>>>(d_state, h->d_logged, h->n_logged, row, h->n);
    }
    NPS_CUDA(cudaGetLastError());
    return 0;
}

static const char* const kMaintActionNames[MA_N_ACTIONS] = {
    "oil_change", "oil_top_off", "bearing_replacement", "seal_replacement", "component_overhaul", "system_cleaning",
    "bearing_inspection", "impeller_inspection", "impeller_replacement", "lubrication_system_check", "motor_inspection",
    "oil_analysis", "vibration_analysis", "tsp_chemical_cleaning", "tsp_mechanical_cleaning", "tube_bundle_inspection",
    "moisture_separator_maintenance", "scale_removal", "eddy_current_testing", "secondary_side_cleaning",
    "routine_maintenance", "tube_interior_scale_cleaning", "primary_scale_cleaning", "cleaning", "blade_replacement",
    "overhaul", "condenser_tube_cleaning", "condenser_tube_plugging", "condenser_chemical_cleaning", "vacuum_system_test",
    "vacuum_leak_detection", "turbine_performance_test", "turbine_system_optimization", "turbine_protection_test",
    "thermal_stress_analysis", "system_coordination_maintenance", "system_steam_quality_maintenance",
    "load_balancing_maintenance", "water_chemistry_adjustment", "tsp_inspection", "tsp_flow_test",
    "tube_interior_inspection", "tube_interior_eddy_current_testing", "primary_chemistry_optimization",
    "condenser_water_treatment", "turbine_oil_change", "turbine_oil_top_off", "oil_filter_replacement",
    "oil_cooler_cleaning", "lubrication_system_test", "vacuum_ejector_cleaning", "vacuum_ejector_nozzle_replacement",
    "vacuum_ejector_inspection", "vacuum_ejector_mechanical_cleaning", "other"};

int nps_n_maintenance_actions(void) { return MA_N_ACTIONS; }
const char* nps_maintenance_action_name(int action) {
    return (action >= 0 && action < MA_N_ACTIONS) ? kMaintActionNames[action] : nullptr;
}

int nps_apply_maintenance(nps_handle* h, double* d_state, const int32_t* h_plant, const int32_t* h_target,
                          const int32_t* h_action, const int32_t* h_arg, int n_requests, int32_t* h_status,
                          void* cuda_stream) {
    if (!h || !d_state || n_requests < 0) return fail("nps_apply_maintenance: bad arguments");
    if (n_requests == 0) return 0;
    if (!h_plant || !h_target || !h_action || !h_status) return fail("nps_apply_maintenance: null request arrays");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    // group requests by plant, keeping the caller's order inside each plant (stable)
    std::vector<int> order(n_requests);
    for (int i = 0; i < n_requests; ++i) {
        if (h_plant[i] < 0 || h_plant[i] >= h->n) return fail("nps_apply_maintenance: plant index out of range");
        if (h_target[i] < 0 || h_target[i] >= MT_N_TARGETS) return fail("nps_apply_maintenance: target out of range");
        if (h_action[i] < 0 || h_action[i] >= MA_N_ACTIONS) return fail("nps_apply_maintenance: action out of range");
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_plant[a] < h_plant[b]; });
    std::vector<MaintRequest> req(n_requests);
    std::vector<int32_t> start;
    for (int j = 0; j < n_requests; ++j) {
        const int i = order[j];
        req[j] = MaintRequest{h_plant[i], h_target[i], h_action[i], h_arg ? h_arg[i] : 0};
        if (j == 0 || req[j].plant != req[j - 1].plant) start.push_back(j);
    }
    const int n_groups = (int)start.size();
    start.push_back(n_requests);
    nps_handle::MaintBuf& mb = h->mb;
    if (mb.cap < n_requests) {
        int cap = mb.cap ? mb.cap : 1024;
        while (cap < n_requests) cap *= 2;
        cudaFree(mb.d_req); cudaFree(mb.d_start); cudaFree(mb.d_status);
        cudaFreeHost(mb.h_req); cudaFreeHost(mb.h_start); cudaFreeHost(mb.h_status);
        mb = nps_handle::MaintBuf();
        NPS_CUDA(cudaMalloc(&mb.d_req, sizeof(MaintRequest) * cap));
        NPS_CUDA(cudaMalloc(&mb.d_start, sizeof(int32_t) * (cap + 1)));
        NPS_CUDA(cudaMalloc(&mb.d_status, sizeof(int32_t) * cap));
        NPS_CUDA(cudaMallocHost(&mb.h_req, sizeof(MaintRequest) * cap));
        NPS_CUDA(cudaMallocHost(&mb.h_start, sizeof(int32_t) * (cap + 1)));
        NPS_CUDA(cudaMallocHost(&mb.h_status, sizeof(int32_t) * cap));
        mb.cap = cap;
    }
    std::memcpy(mb.h_req, req.data(), sizeof(MaintRequest) * n_requests);
    std::memcpy(mb.h_start, start.data(), sizeof(int32_t) * start.size());
    NPS_CUDA(cudaMemcpyAsync(mb.d_req, mb.h_req, sizeof(MaintRequest) * n_requests, cudaMemcpyHostToDevice, s));
    NPS_CUDA(cudaMemcpyAsync(mb.d_start, mb.h_start, sizeof(int32_t) * start.size(), cudaMemcpyHostToDevice, s));
    const int block = 32;
    nps_maintenance_kernel<<<(n_groups + block - 1) / block, block, 0, s>>>(d_state, h->params, (const MaintRequest*)mb.d_req, mb.d_start,
                                                                           n_groups, mb.d_status, h->n);
    NPS_CUDA(cudaGetLastError());
    NPS_CUDA(cudaMemcpyAsync(mb.h_status, mb.d_status, sizeof(int32_t) * n_requests, cudaMemcpyDeviceToHost, s));
    NPS_CUDA(cudaStreamSynchronize(s));
    for (int j = 0; j < n_requests; ++j) h_status[order[j]] = mb.h_status[j];
    return 0;
}

int nps_read_fields(nps_handle* h, const double* d_state, const int32_t* fields, int n_fields, double* out_host,
                    void* cuda_stream) {
    if (!h || !d_state || !fields || !out_host || n_fields <= 0) return fail("nps_read_fields: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    for (int i = 0; i < n_fields; ++i) if (fields[i] < 0 || fields[i] >= kNState) return fail("nps_read_fields: field index out of range");
    if (h->gather_cap < n_fields) {
        cudaFree(h->d_gather_fields); cudaFree(h->d_gather_out);
        h->d_gather_fields = nullptr; h->d_gather_out = nullptr; h->gather_cap = 0;
        NPS_CUDA(cudaMalloc(&h->d_gather_fields, sizeof(int32_t) * n_fields));
        NPS_CUDA(cudaMalloc(&h->d_gather_out, sizeof(double) * n_fields * h->n));
        h->gather_cap = n_fields;
    }
    // everything on the caller's stream: ordered after the launches already queued there, no legacy-default-stream work
    NPS_CUDA(cudaMemcpyAsync(h->d_gather_fields, fields, sizeof(int32_t) * n_fields, cudaMemcpyHostToDevice, s));
    const int block = 128;
    nps_gather_kernel<<<(int)((h->n + block - 1) / block), block, 0, s>>>(d_state, h->d_gather_fields, n_fields, h->d_gather_out, h->n);
    NPS_CUDA(cudaGetLastError());
    NPS_CUDA(cudaMemcpyAsync(out_host, h->d_gather_out, sizeof(double) * n_fields * h->n, cudaMemcpyDeviceToHost, s));
    NPS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
