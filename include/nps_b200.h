/* nps_b200 — C ABI of the B200-native batched plant-dynamics engine.
 *
 * Drop-in boundary for the per-timestep hot path of NuclearnAI/nuclear-sim:
 *   NuclearPlantSimulator.step()            nuclear_simulator/simulator/core/sim.py:130-258
 *   get_observation() / calculate_reward()  nuclear_simulator/simulator/core/sim.py:290-333, 500-544
 *   StateManager threshold monitoring       nuclear_simulator/simulator/state/state_manager.py:1307-1369
 *   StateManager row collection / export    nuclear_simulator/simulator/state/state_manager.py:152-233
 * The reference has no FFI of its own (it is pure Python); these are the entry points a ctypes
 * binding on the reference side would call.
 *
 * Conventions
 *   - plain pointers and sizes only; the CALLER owns every buffer, the library borrows it for the call;
 *   - plant state is one FP64 structure-of-arrays slab in device memory: field f of plant p is
 *     slab[f * n_plants + p]; field order = struct PlantState in nuclear-sim_b200/csrc/plant/state.h
 *     (nps_field_name(i) returns the flat name of field i);
 *   - device entry points are asynchronous on the given CUDA stream (cudaStream_t passed as void*),
 *     no hidden synchronisation; *_host entry points take host buffers and synchronise;
 *   - return 0 on success, negative on error (nps_last_error() gives the message);
 *   - one handle per device per process; a handle is not thread-safe.
 */
#ifndef NPS_B200_H
#define NPS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nps_handle nps_handle;

#define NPS_ABI_VERSION 2
#define NPS_OBS_DIM 22          /* get_observation(): 12 primary + 6 secondary + 4 feedwater (sim.py:290-333) */
#define NPS_NOISE_PER_STEP 5    /* z_heat, z_ph, u_ph[3] (constant_heat_source.py:178; ph_control_system.py:278-420) */

int nps_abi_version(void);
const char* nps_last_error(void);

/* layout introspection (PlantState / PlantParams of csrc/plant/state.h) */
int nps_n_state(void);
int nps_n_params(void);
const char* nps_field_name(int field);        /* e.g. "fw.pump[0].lub.oil_level" */
const char* nps_param_name(int param);

/* lifetime: replaces NuclearPlantSimulator.__init__ for n_plants plants (sim.py:30-87) */
int nps_create(int64_t n_plants, int device, nps_handle** out);
void nps_destroy(nps_handle* h);
int64_t nps_n_plants(const nps_handle* h);

/* batch-uniform parameters (config dataclasses of systems/secondary/<subsystem>/config.py); host array of nps_n_params() */
int nps_set_params(nps_handle* h, const double* params_host, int n_params);

/* One launch = k_substeps calls of NuclearPlantSimulator.step() for every plant (sim.py:130-258).
 *   d_state      [n_state][n_plants]            in/out
 *   d_action     [k_substeps][n_plants] int8    ControlAction values (primary/__init__.py:28-45), may be NULL = NO_ACTION
 *   d_magnitude  [k_substeps][n_plants]         may be NULL = 1.0
 *   d_noise      [k_substeps][5][n_plants]      host-supplied random streams, may be NULL = (0,0,1,1,1)
 *   d_setpoint   [k_substeps][n_plants]         heat_source.set_power_setpoint(%) applied before each substep
 *                                               (maintenance_scenario_runner.py:651-671); NaN entries / NULL = unchanged
 *   d_obs        [22][n_plants]   observation after the last substep, may be NULL
 *   d_reward     [n_plants]       calculate_reward after the last substep, may be NULL
 *   d_done       [n_plants] uint8 1 if a scram was activated in any substep of this launch, may be NULL */
int nps_step(nps_handle* h, double* d_state, const int8_t* d_action, const double* d_magnitude,
             const double* d_noise, const double* d_setpoint, int k_substeps, double* d_obs, double* d_reward,
             uint8_t* d_done, void* cuda_stream);

/* --- per-step monitoring inside a fused launch ---------------------------------------------------------------
 * The reference looks at the plant after EVERY step: `done` = scram_activated (sim.py:256), the NaN reset
 * (thermal_hydraulics.py:257-269), threshold monitoring (state_manager.py:1307-1369).  nps_step_monitored does the same
 * after every substep of a K-substep launch, so the step of every discrete event is the one K = 1 launches would report.
 * All pointers are caller-owned device memory; any of them may be NULL (that output is skipped).
 *   d_last_fired   [n_thresholds][n_plants] cooldown stamps (same array nps_check_thresholds uses).  Non-NULL switches on
 *                  threshold evaluation after every substep (rows from nps_set_thresholds); violations are appended to
 *                  d_events[*d_n_events ...] (one warp-aggregated atomic per row and warp), *d_n_events keeps counting
 *                  past event_capacity so the caller can detect an overflow; the caller zeroes it.
 *   skip_last_check  1: no threshold evaluation after the LAST substep (the caller applies maintenance first and then
 *                  runs nps_check_thresholds for that step, as sim.py:209-223 orders it).
 *   d_watch_fields [n_watch] (<= 32) PlantState field indices; d_watch_step [n_watch][n_plants] = step index at which the
 *                  field first was non-zero, -1 = not yet (caller initialises to -1; persists across launches).
 *   d_first_scram_step, d_first_nan_reset_step [n_plants] int32, -1 = never; d_status [n_plants] sticky NPS_STATUS_* bits.
 *   d_reward_k [k_substeps][n_plants], d_done_k [k_substeps][n_plants] uint8: calculate_reward / done after every substep.
 *   step0          step index of the first substep of this launch (event.step = step0 + substep). */
#define NPS_STATUS_NAN_RESET 1u
#define NPS_STATUS_SCRAM 2u
typedef struct nps_event { int32_t plant; int32_t row; int32_t step; int32_t reserved; double value; double time_minutes; } nps_event;
typedef struct nps_monitor {
    double* d_last_fired;
    nps_event* d_events;
    uint32_t* d_n_events;
    uint32_t event_capacity;
    int32_t skip_last_check;
    const int32_t* d_watch_fields;
    int32_t* d_watch_step;
    int32_t n_watch;
    int32_t reserved;
    int32_t* d_first_scram_step;
    int32_t* d_first_nan_reset_step;
    uint32_t* d_status;
    double* d_reward_k;
    uint8_t* d_done_k;
    int64_t step0;
} nps_monitor;
int nps_step_monitored(nps_handle* h, double* d_state, const int8_t* d_action, const double* d_magnitude,
                       const double* d_noise, const double* d_setpoint, int k_substeps, double* d_obs, double* d_reward,
                       uint8_t* d_done, const nps_monitor* monitor, void* cuda_stream);

/* Device-side noise.  The reference draws its per-step noise inside its Python objects (constant_heat_source.py:178,
 * ph_control_system.py:288,409-420); parity runs pass those very streams in d_noise.  A production batch that passes
 * d_noise == NULL can instead have every (plant, step) draw its five numbers on the device from Philox4x32-10 keyed by
 * `seed` (counter = (plant_offset + plant, step)): reproducible, independent of batch shape, k_substeps and the number
 * of GPUs when plant_offset is the rank's first global plant id.  `first_step` is the step index of the next launch;
 * every launch advances it by k_substeps.  enabled = 0 restores "no noise array = no noise".
 * nps_device_rng_draws evaluates the same generator on the host (z_heat, z_ph, u0, u1, u2). */
int nps_set_device_rng(nps_handle* h, int enabled, uint64_t seed, uint64_t plant_offset, uint64_t first_step);
int nps_device_rng_draws(uint64_t seed, uint64_t plant, uint64_t step, double* out5);

/* Test hook: out[i] = x[i] ** y[i] through the step kernel's own power function (csrc/plant/hd.h py_pow: exact
 * specialisations, csrc/plant/fastpow.h for positive finite bases, libdevice pow otherwise), device arrays. */
int nps_selftest_pow(const double* d_x, const double* d_y, double* d_out, int64_t n, void* cuda_stream);

/* Launch shape of nps_step for batches up to 148 x 4 x 32 = 18,944 plants (larger batches always run one thread per plant):
 *   0 (default)  two threads per plant — one advances primary side / feedwater / steam generators / chemistry, the other
 *                the turbine and condenser one substep behind it (they are pure sinks of the step's dataflow), so the
 *                dependency chain per substep is the longer half instead of the sum;
 *   1            one thread per plant.
 * Both produce bit-identical state; the environment variable NPS_SMALL_SHAPE=1 selects shape 1 at nps_create. */
int nps_set_small_batch_shape(nps_handle* h, int shape);

/* Measurement hook: achieved FP64 FMA throughput of `device` (8 independent DFMA chains per thread, 2 048 threads per
 * SM, best of 4 timed launches of `iters` iterations; synchronous).  The FP64 side of the roofline in bench.py divides by
 * this number instead of a datasheet figure. */
int nps_measure_fp64_peak(int device, int iters, double* out_tflops, double* out_ms);

/* Same step with HOST buffers for the per-step inputs and outputs (pinned or pageable): the
 * host->device copies of action/magnitude/noise and the device->host copies of obs/reward/done are
 * issued on the stream inside the call, which returns after they complete. State stays on device. */
int nps_step_host(nps_handle* h, double* d_state, const int8_t* h_action, const double* h_magnitude,
                  const double* h_noise, const double* h_setpoint, int k_substeps, double* h_obs, double* h_reward,
                  uint8_t* h_done, void* cuda_stream);

/* Pipelined form of nps_step_host for a driver that steps in a loop: returns at once with a ticket in
 * [0, NPS_PIPE_DEPTH); the host->device copies run on an internal copy stream into one of NPS_PIPE_DEPTH staging sets,
 * so the inputs of later launches travel while launch i computes, and the results of launch i travel while launch i+1
 * computes.  Host buffers must be pinned and must stay untouched until nps_wait(ticket) returns (inputs) / are valid
 * after it returns (outputs).  At most NPS_PIPE_DEPTH launches may be outstanding: tickets are handed out round-robin,
 * so wait for the ticket of call i - NPS_PIPE_DEPTH before issuing call i. */
#define NPS_PIPE_DEPTH 4
int nps_pipe_depth(void);
int nps_step_host_async(nps_handle* h, double* d_state, const int8_t* h_action, const double* h_magnitude,
                        const double* h_noise, const double* h_setpoint, int k_substeps, double* h_obs, double* h_reward,
                        uint8_t* h_done, void* cuda_stream);
int nps_wait(nps_handle* h, int ticket);

/* get_observation()/calculate_reward() of the current state without stepping */
int nps_observe(nps_handle* h, const double* d_state, double* d_obs, double* d_reward, void* cuda_stream);

/* --- threshold monitoring (StateManager._check_maintenance_thresholds, state_manager.py:1307-1369) ---
 * A threshold row is (field, comparator, value, cooldown_minutes); comparator: 0 '>', 1 '<', 2 '>=', 3 '<=',
 * 4 '=='(|v-x| < 1e-3), 5 '!='.  field >= 0: PlantState field; field == -1: inert row (the reference finds no logged
 * column for it, state_manager.py:1371-1410); field <= -2: derived column, code -(2 + 4*kind + unit), kind 0 =
 * sum_wear_level of feedwater pump `unit` (feedwater/pump_lubrication.py:1585-1596).
 * d_last_fired [n_thresholds][n_plants] holds the time (minutes) each threshold last fired (-inf = never).
 * Output: d_flags [n_words][n_plants] uint32 bit t%32 of word t/32 set when threshold t fired at time
 * now; d_any_warp [ceil(n_plants/32)] uint32 = warp ballot of "plant has any flag" (host drain index). */
int nps_set_thresholds(nps_handle* h, const int32_t* field, const int32_t* comparator, const double* value,
                       const double* cooldown_minutes, int n_thresholds);
int nps_check_thresholds(nps_handle* h, const double* d_state, double* d_last_fired, uint32_t* d_flags,
                         uint32_t* d_any_warp, void* cuda_stream);

/* The same check with the violations appended to an event list (struct nps_event, see nps_step_monitored; event.step =
 * `step`): the form the work-order bookkeeping consumes, identical for checks inside a fused launch and between launches. */
int nps_check_thresholds_events(nps_handle* h, const double* d_state, double* d_last_fired, nps_event* d_events,
                                uint32_t* d_n_events, uint32_t event_capacity, int32_t step, void* cuda_stream);

/* --- trajectory ring buffer (StateManager.collect_states/_add_row, state_manager.py:152-233) ---
 * Appends the selected fields of every plant as row (write_index % ring_rows) of
 * d_ring [ring_rows][n_logged][n_plants].  Each logged field is one contiguous row of the slab; the rows travel as
 * cp.async.bulk (TMA) transfers global -> shared -> global in a 4-stage pipeline (0.90 of the HBM copy peak at 65,536
 * plants).  Plant counts that are odd (rows not 16-byte aligned) take a shared-memory tile kernel instead. */
int nps_set_logged_fields(nps_handle* h, const int32_t* fields, int n_logged);
int nps_log_row(nps_handle* h, const double* d_state, double* d_ring, int64_t ring_rows, int64_t write_index,
                void* cuda_stream);

/* --- maintenance action effects (AutoMaintenanceSystem._execute_work_order -> component.perform_maintenance,
 *     systems/maintenance/auto_maintenance.py:504-673) ---
 * Applies n_requests (plant, target, action, arg) requests IN ORDER to the device state; requests for the same plant
 * are serialised.  target: 0-3 FWP-1..4, 4 FEE-001, 5-7 SG-0..2, 8 SG system, 9-22 HP-1..LP-6, 23 turbine, 24 condenser,
 * 25 TB-LUB-001 (turbine bearing lubrication), 26-27 SJE-001..2 (steam jet ejectors).
 * action: index into nps_maintenance_action_name(); arg: bearing selector for bearing_replacement
 * (0 all, 1 motor_bearings, 2 pump_bearings, 3 thrust_bearing).  h_status[i]: 0 failed (MaintenanceResult.success
 * False), 1 success, 2 target has no restated perform_maintenance.  Host arrays; synchronises on the stream. */
int nps_n_maintenance_actions(void);
const char* nps_maintenance_action_name(int action);
int nps_apply_maintenance(nps_handle* h, double* d_state, const int32_t* h_plant, const int32_t* h_target,
                          const int32_t* h_action, const int32_t* h_arg, int n_requests, int32_t* h_status,
                          void* cuda_stream);

/* gather selected fields of all plants to a host array out[n_fields][n_plants]; runs on `cuda_stream` (ordered after
 * the launches queued there) and returns after synchronising that stream */
int nps_read_fields(nps_handle* h, const double* d_state, const int32_t* fields, int n_fields, double* out_host,
                    void* cuda_stream);

/* --- work-order table: the host side of the automatic maintenance loop, native (no device work) ---
 * What AutoMaintenanceSystem keeps per plant, for N plants: StateManager._check_maintenance_thresholds batching
 * (simulator/state/state_manager.py:1307-1369: the violations of one component in one step are ONE event),
 * _create_automatic_work_order (systems/maintenance/auto_maintenance.py:398-466: known action, the 24-"hour" dedupe stamp
 * compared in minutes, no active order for the same (component, action), priority delay) and update /
 * _execute_work_order (auto_maintenance.py:200-236, 504-580: due orders in creation order, at HEAD one per plant per
 * update).  Threshold rows are described once: component of the row, its rule table (value > rule_thr[r][j] selects
 * rule_act[r][j], first match wins), fallback action, own action, priority 0..4 (LOW..EMERGENCY), sub-component.
 * All arrays are host arrays owned by the caller; every function returns < 0 on bad arguments. */
typedef struct nps_wo_table nps_wo_table;
int nps_wo_create(int64_t n_plants, int n_components, int n_rows, int max_rules, const int64_t* row_comp,
                  const double* rule_thr, const int64_t* rule_act, const int64_t* fallback, const int64_t* row_action,
                  const int64_t* row_prio, const int64_t* row_sub, const double* prio_delay_minutes, double dedupe_window,
                  int head_quirks, nps_wo_table** out);
void nps_wo_destroy(nps_wo_table* t);
/* violations of one step, sorted by (plant, row) -> groups per (plant, component) with the single-violation decision;
 * groups with several violations are counted in *n_multi (MaintenanceOrchestrator decides those: the caller patches
 * g_act / g_prio / g_sub before nps_wo_issue).  Returns the number of groups. */
int64_t nps_wo_group(const nps_wo_table* t, int64_t n, const int64_t* plant, const int64_t* row, const double* value,
                     int64_t* g_start, int64_t* g_count, int64_t* g_plant, int64_t* g_comp, int64_t* g_act,
                     int64_t* g_prio, int64_t* g_sub, int64_t* n_multi);
/* decided events -> work orders; out_group / out_seq: group index and WO number of every order created; returns how many */
int64_t nps_wo_issue(nps_wo_table* t, double t_minutes, int64_t n_groups, const int64_t* g_plant, const int64_t* g_comp,
                     const int64_t* g_act, const int64_t* g_prio, const int64_t* g_sub, const uint8_t* act_ok,
                     int64_t n_actions, int64_t* out_group, int64_t* out_seq);
int64_t nps_wo_n_pending(const nps_wo_table* t);
/* the orders update(t_minutes) executes, as columns (capacity cap each); returns the count, or the capacity needed */
int64_t nps_wo_due(nps_wo_table* t, double t_minutes, int64_t cap, int64_t* plant, int64_t* comp, int64_t* act,
                   int64_t* prio, int64_t* sub, int64_t* seq, double* created, double* planned);
int nps_wo_complete(nps_wo_table* t);                                    /* the orders of the last nps_wo_due are done */
int nps_wo_reset_plants(nps_wo_table* t, const int64_t* plants, int64_t n);
int nps_wo_sizes(const nps_wo_table* t, int64_t* n_pending, int64_t* n_stamps);
int nps_wo_export(const nps_wo_table* t, int64_t* pend_cols, double* pend_times, int64_t* stamp_keys, double* stamp_times,
                  int64_t* n_created);
int nps_wo_import(nps_wo_table* t, int64_t n_pending, const int64_t* pend_cols, const double* pend_times, int64_t n_stamps,
                  const int64_t* stamp_keys, const double* stamp_times, const int64_t* n_created);

#ifdef __cplusplus
}
#endif
#endif
