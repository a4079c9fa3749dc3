/* tsc_b200.h -- C ABI of the B200-native batched traffic-signal-control engine.
 *
 * This library replaces, for B scenario replicas at once, the one native object
 * the reference's CityFlow backend drives: `cityflow.Engine` (third-party
 * C++/pybind11, not vendored by the reference).  Each entry point cites the
 * reference interface it stands in for (paths relative to the reference repo).
 *
 *   cityflow.Engine(config_file, thread_num)   pytsc/backends/cityflow/simulator.py:71-74   -> tsc_create
 *   engine.reset()                             simulator.py:95                              -> tsc_reset
 *   engine.set_tl_phase(id, phase)             backends/cityflow/traffic_signal.py:31,58    -> tsc_set_phase
 *   engine.next_step()  (x delta_time)         simulator.py:76-77,86-88                     -> tsc_step
 *   engine.get_lane_waiting_vehicle_count()    retriever.py:95   \
 *   engine.get_lane_vehicles()                 retriever.py:96    |
 *   engine.get_vehicle_speed()                 retriever.py:97    |  one call, batched      -> tsc_retrieve
 *   engine.get_vehicle_info(v)                 retriever.py:35    |
 *   engine.get_vehicle_count()                 retriever.py:109   |
 *   engine.get_average_travel_time()           retriever.py:110   |
 *   engine.get_current_time()                  simulator.py:50, retriever.py:111           /
 *   TrafficSignalNetwork.step(actions)         pytsc/__init__.py:178-182  (fused fast path) -> tsc_env_step
 *   {FixedTime,Greedy,MaxPressure,SOTL,Random}Controller.get_action
 *                                              pytsc/controllers/controllers.py:39-268      -> tsc_env_step(controller=...),
 *                                                                                              tsc_controller_act
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types; `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).
 *   - every function returns 0 on success, a negative TSC_E* code otherwise;
 *     tsc_last_error() gives the message.  No exceptions cross the ABI.
 *   - the handle owns all simulation state on one device.  Callers own input /
 *     output buffers; pointers documented "device" must be device memory of the
 *     handle's device, "host" pointers are ordinary host memory.
 *   - nothing synchronises the stream except tsc_snapshot, tsc_check and the
 *     *_host convenience calls.  One handle must not be used from two threads
 *     at once.
 *   - device-side capacity overflows / ordering violations set a sticky
 *     per-replica flag that tsc_check reports.
 */
#ifndef TSC_B200_H
#define TSC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSC_ABI_VERSION 3

enum {
    TSC_OK = 0,
    TSC_EINVAL = -1,      /* bad argument / scenario table */
    TSC_ECUDA = -2,       /* CUDA runtime error (message has the detail) */
    TSC_ENOMEM = -3,      /* out of device memory for the replica state */
    TSC_EOVERFLOW = -4,   /* a replica exceeded vehicle_capacity (sticky flag, see tsc_check) */
    TSC_EORDER = -5       /* a vehicle left its drivable out of FIFO order (sticky flag) */
};

/* vehicle template row layout (doubles) */
enum {
    TSC_T_LEN = 0, TSC_T_MAX_POS_ACC, TSC_T_MAX_NEG_ACC, TSC_T_USUAL_POS_ACC, TSC_T_USUAL_NEG_ACC,
    TSC_T_MIN_GAP, TSC_T_MAX_SPEED, TSC_T_HEADWAY, TSC_T_YIELD_DIST, TSC_T_TURN_SPEED, TSC_T_APPROACH_DIST,
    TSC_T_STRIDE = 12
};

enum { TSC_REWARD_QUEUE = 0, TSC_REWARD_PRESSURE = 1 };
enum { TSC_OBS_LANE_FEATURES = 0, TSC_OBS_POSITION_MATRIX = 1 };
enum { TSC_ACT_PHASE_SELECTION = 0, TSC_ACT_PHASE_SWITCH = 1 };
/* who picks the next phase in tsc_env_step: the caller's actions interpreted by the scenario's action
 * space; the in-kernel fixed-time controller; or the caller's actions taken as pytsc phase indices
 * (TSController.switch_phase, backends/cityflow/traffic_signal.py:51-59) whatever the action space. */
enum { TSC_CTRL_EXTERNAL = 0, TSC_CTRL_FIXED_TIME = 1, TSC_CTRL_PHASE_INDEX = 2,
       /* pytsc's rule-based controllers evaluated on the device from the position-matrix windows of the
        * current state (controllers/controllers.py): they return pytsc phase indices, applied like
        * TSC_CTRL_PHASE_INDEX.  controller_arg: GREEDY / MAX_PRESSURE / RANDOM = seed of the counter-based
        * generator that breaks ties (the reference draws np.random.choice among tied actions);
        * SOTL = TSC_SOTL_ARG(theta, mu, phi_min) (reference defaults 3, 4, 5). */
       TSC_CTRL_GREEDY = 3, TSC_CTRL_MAX_PRESSURE = 4, TSC_CTRL_SOTL = 5, TSC_CTRL_RANDOM = 6 };
#define TSC_SOTL_ARG(theta, mu, phi_min) (((theta) & 0xFF) | (((mu) & 0xFF) << 8) | (((phi_min) & 0xFFFF) << 16))
/* score written by tsc_controller_act for an action the mask forbids (the reference's float("-inf")) */
#define TSC_SCORE_MASKED INT32_MIN

/* A compiled scenario: flat, read-only tables built on the host by
 * pytsc_b200.scenario.compile_scenario().  All pointers are HOST pointers,
 * borrowed for the duration of tsc_create only (the library copies them to the
 * device).  Drivable index d: lanes 0..n_lanes-1 (roadnet road order, lane
 * order), lane-links n_lanes..n_lanes+n_lanelinks-1 (intersection order,
 * road-link order, lane-link order).  Signals ("agents") are the non-virtual
 * intersections in roadnet order -- the order of pytsc's traffic_signals dict
 * (backends/cityflow/network_parser.py:49-78). */
typedef struct tsc_scenario {
    int32_t abi_version;          /* TSC_ABI_VERSION */
    int32_t n_lanes, n_lanelinks, n_signals;
    int32_t n_vehicles;           /* spawn list length N (vehicles the flows create within horizon_ticks) */
    int32_t n_templates;
    int32_t n_route_seq;          /* length of route_seq */
    int32_t n_cross_entries;      /* length of xr_* (two per cross) */
    int32_t horizon_ticks;        /* ticks the spawn list covers (sim_length + initial_wait_time) */
    int32_t max_raw_phases;       /* row stride of sig_phase_mask */
    int32_t max_phases;           /* P: row stride of the pytsc phase tables */
    int32_t n_in_total, n_out_total, n_nbr_total;
    int32_t n_ctl_total;          /* length of ctl_in_lane / ctl_out_lane */
    int32_t n_dm_total;           /* length of dm_lane */
    int32_t n_flow_sets;          /* F >= 1: alternative flow files compiled into this scenario (pytsc's
                                     cityflow.flow_files with flow_rate_type random / sequential,
                                     backends/cityflow/config.py:63-76; DisruptedConfig :106-175).  Every replica
                                     runs one of them, chosen at reset (tsc_reset_flows). */

    /* --- engine tables (CityFlow semantics) --- */
    const double  *drv_length;        /* [D] */
    const double  *drv_max_speed;     /* [D] lane-links: 10000 */
    const int32_t *lane_ll_off;       /* [L+1] CSR: lane -> lane-links leaving it (roadnet appearance order) */
    const int32_t *lane_ll;           /* lane-link index 0..K-1 */
    const int32_t *lane_spawn_off;    /* [F][L+1] CSR per flow set: lane -> vehicles that start on it, FIFO order;
                                         rows index one shared lane_spawn_vid (row f starts where row f-1 ended) */
    const int32_t *lane_spawn_vid;    /* [N] vehicle ids (the N vehicles of all flow sets, set-major) */
    const int32_t *ll_start_lane;     /* [K] */
    const int32_t *ll_end_lane;       /* [K] */
    const int32_t *ll_signal;         /* [K] signal (agent) index */
    const int32_t *ll_roadlink;       /* [K] road-link index inside the intersection (bit of the phase mask) */
    const int32_t *ll_type;           /* [K] 3 go_straight, 2 turn_left, 1 turn_right */
    const int32_t *ll_cross_off;      /* [K+1] CSR: lane-link -> crosses, ascending distance along the link */
    const double  *xr_dist;           /* distance of the cross along this lane-link */
    const int32_t *xr_foe_ll;         /* the other lane-link of the cross */
    const double  *xr_foe_dist;       /* distance of the cross along the other lane-link */
    const uint32_t *sig_phase_mask;   /* [A][max_raw_phases] bit r = road-link r available */
    const int32_t *sig_n_raw_phases;  /* [A] */
    const int32_t *route_seq;         /* -1, d0, d1, ..., -1, d0, ... drivable sequences, -1 separated, leading -1 */
    const int32_t *veh_tick;          /* [N] tick at which the flow creates the vehicle (creation order) */
    const int32_t *veh_seq_start;     /* [N] index into route_seq of the vehicle's first drivable */
    const int32_t *veh_tmpl;          /* [N] */
    const int32_t *veh_priority;      /* [N] mt19937 draw (yield tie-break only) */
    const double  *tmpl;              /* [T][TSC_T_STRIDE] */

    /* --- pytsc tables (Retriever / TrafficSignal / reward / mask / observation) --- */
    const double  *lane_pytsc_length; /* [L] centre-to-centre road length (network_parser.py:325-352) */
    const double  *lane_feat;         /* [L][9] static lane features (observations.py:90-116) */
    const int32_t *sig_in_off;        /* [A+1] CSR incoming lanes, pytsc (sorted id) order */
    const int32_t *sig_in_lane;
    const int32_t *sig_out_off;       /* [A+1] CSR outgoing lanes */
    const int32_t *sig_out_lane;
    const int32_t *sig_n_phases;      /* [A] number of pytsc phases */
    const int32_t *sig_phase_raw;     /* [A][P] raw light-phase of pytsc phase p */
    const uint8_t *sig_phase_green;   /* [A][P] 1 = green index */
    const int32_t *sig_min_time;      /* [A][P] */
    const int32_t *sig_max_time;      /* [A][P] */
    const int32_t *nbr_off;           /* [A+1] CSR: reward neighbours in the reference's summation order */
    const int32_t *nbr_idx;
    const double  *nbr_weight;        /* gamma**k */
    /* rule-based controllers: phase_to_inc_out_lanes (backends/cityflow/network_parser.py:212-257) of
     * pytsc phase p of signal s = entries ctl_off[s*P+p] .. ctl_off[s*P+p+1] */
    const int32_t *ctl_off;           /* [A*P+1] */
    const int32_t *ctl_in_lane;       /* incoming lane of the entry */
    const int32_t *ctl_out_lane;      /* LAST outgoing lane listed for it (what controllers.py:165-169 ends up using), -1 = none */
    /* network-level graph metrics (MetricsParser.density_map, backends/cityflow/metrics.py:170-199): the lanes leading
     * from signal i to signal j (parsed_network.neighbors_lanes, agent order) = entries dm_off[i*A+j] .. dm_off[i*A+j+1] */
    const int32_t *dm_off;            /* [A*A+1] */
    const int32_t *dm_lane;
    const double  *dm_adjacency;      /* [A][A] parsed_network.adjacency_matrix as the reference adds it (its own, sorted-id, order) */

    /* --- options --- */
    int32_t reward_type, obs_type, action_space, round_robin;
    int32_t visibility;               /* bins (signal.visibility) */
    int32_t yellow_time;              /* == delta_time */
    int32_t obs_dim, state_dim, n_actions;
    int32_t reference_exact;          /* reproduce pad_list int truncation (common/utils.py:91-112) */
    int32_t max_lanes_per_signal;     /* 16 (observations.py max_n_controlled_lanes) */
    int32_t max_obs_phases;           /* 20 */
    double  veh_size_min_gap;         /* 7.5 */
    double  flickering_coef;
    double  interval;                 /* 1.0 */
} tsc_scenario_t;

/* Output buffers of tsc_retrieve / tsc_env_step.  DEVICE pointers owned by the
 * caller; any may be NULL to skip that output.  Shapes in comments; B-major,
 * contiguous. */
typedef struct tsc_outputs {
    int32_t *lane_count;        /* [B][L]  vehicles on lane            (retriever.py:66,78) */
    int32_t *lane_queued;       /* [B][L]  of those, speed < 0.1       (retriever.py:64,95) */
    float   *lane_occupancy;    /* [B][L]                               (retriever.py:74-76) */
    float   *lane_mean_speed;   /* [B][L]                               (retriever.py:67-73) */
    double  *lane_meas64;       /* [B][L][2] occupancy, mean_speed in fp64 (compatibility view) */
    float   *pos_in;            /* [B][n_in_total][visibility]  last `visibility` bins   (traffic_signal.py:124) */
    float   *pos_out;           /* [B][n_out_total][visibility] first `visibility` bins  (traffic_signal.py:135) */
    double  *sig_stats64;       /* [B][A][8] n_queued, occupancy, mean_speed, mean_delay, out_occupancy,
                                             pressure, norm_time_on_phase, phase_index   (traffic_signal.py:101-141) */
    float   *obs;               /* [B][A][obs_dim]    (observations.py:140-160 | 305-329) */
    float   *state;             /* [B][A][state_dim]  (observations.py:352-374) */
    float   *reward;            /* [B][A] local rewards (reward.py:67-88 | 115-136) */
    float   *reward_global;     /* [B]    (reward.py:54-65 | 102-113) */
    uint8_t *mask;              /* [B][A][n_actions]  (actions.py:119-131 | 169-188) */
    double  *sim;               /* [B][4] n_vehicles, average_travel_time, time_step, n_finished (retriever.py:101-112) */
    double  *metrics;           /* [B][8] n_queued, mean_speed, mean_delay, density, pressure, network_flow,
                                          flickering_signal, norm_mean_speed  (backends/cityflow/metrics.py:221-232) */
    double  *density_map;       /* [B][A][A] MetricsParser.density_map: clip(mean occupancy of the lanes i -> j, 0, 1), symmetrised,
                                       + 1e-6 * adjacency  (backends/cityflow/metrics.py:170-199) */
    int32_t *err;               /* [B] sticky error bits of the replica (0 = healthy; 1 | 4 vehicle capacity exceeded,
                                       2 FIFO order lost, 8 bad light phase): a replica with a bit set is frozen and its
                                       rows are stale -- the same bits tsc_check reports, readable without a sync */
} tsc_outputs_t;

typedef struct tsc_engine *tsc_handle;

/* Library / build information. */
int  tsc_abi_version(void);
const char *tsc_last_error(void);

/* Allocate B replicas of `scn` on `device`.  vehicle_capacity = upper bound on
 * simultaneously running vehicles per replica (slots), 0 = library default. */
int  tsc_create(const tsc_scenario_t *scn, int32_t n_replicas, int32_t device,
                int32_t vehicle_capacity, tsc_handle *out);
void tsc_destroy(tsc_handle h);

/* Sizes the caller needs to allocate outputs. */
int  tsc_get_dims(tsc_handle h, int32_t *n_replicas, int32_t *n_lanes, int32_t *n_signals,
                  int32_t *obs_dim, int32_t *state_dim, int32_t *n_actions,
                  int32_t *n_in_total, int32_t *n_out_total, int32_t *visibility);

/* All replicas back to tick 0, empty network, light phase 0, program state cleared; every replica keeps
 * the flow set it was last given (0 after tsc_create). */
int  tsc_reset(tsc_handle h, void *stream);

/* The same, and replica b restarts on flow set flow_set_per_replica[b] (host int32 [B], each in
 * [0, n_flow_sets)): what a new cityflow.Engine built on a newly drawn flow file is to the reference
 * (pytsc/__init__.py:164-176 -> backends/cityflow/config.py:63-76, per replica).  NULL = keep. */
int  tsc_reset_flows(tsc_handle h, const int32_t *flow_set_per_replica, void *stream);

/* Selected replicas back to tick 0 (as tsc_reset does for all of them), e.g. the replicas whose
 * episode has ended while the others run on.  replicas: host int32 [n], indices in [0, B).  The
 * caller re-applies tsc_init_program / phases as after tsc_reset.  Ordered on `stream`. */
int  tsc_reset_replicas(tsc_handle h, const int32_t *replicas, int32_t n, void *stream);
/* The same with a new flow set for each of them (host int32 [n]; NULL = keep). */
int  tsc_reset_replicas_flows(tsc_handle h, const int32_t *replicas, const int32_t *flow_sets, int32_t n, void *stream);

/* Engine state snapshot / restore -- the batched counterpart of CityFlow's engine.snapshot() /
 * engine.load(archive) (cityflow.Engine API; pytsc's save_replay / mid-episode restarts).
 * tsc_state_bytes: size of the opaque blob holding all B replicas (vehicles, waiting-buffer cursors,
 * signal programs, travel-time sums, tick).  tsc_save_state / tsc_load_state copy it to / from `buf`,
 * host or device memory of at least that size; ordered on `stream`, which is synchronised before
 * returning.  A blob is valid for handles created from the same scenario with the same
 * n_replicas and vehicle_capacity; tsc_load_state checks the header it wrote. */
int64_t tsc_state_bytes(tsc_handle h);
int  tsc_save_state(tsc_handle h, void *buf, int64_t buf_bytes, void *stream);
int  tsc_load_state(tsc_handle h, const void *buf, int64_t buf_bytes, void *stream);

/* raw_phase: device int32 [B][A] raw CityFlow light-phase index per signal. */
int  tsc_set_phase(tsc_handle h, const int32_t *raw_phase, void *stream);

/* Initialise pytsc's per-signal program (TSProgram.set_initial_phase,
 * common/traffic_signal.py:83-92; backends/cityflow/traffic_signal.py:26-32):
 * current phase index = phase_index, time_on_phase = 0, and the raw phase set. */
int  tsc_init_program(tsc_handle h, int32_t phase_index, void *stream);

/* n_ticks x engine.next_step() on every replica. */
int  tsc_step(tsc_handle h, int32_t n_ticks, void *stream);

/* Retriever + per-signal stats + reward / mask / observation from the current state. */
int  tsc_retrieve(tsc_handle h, const tsc_outputs_t *out, void *stream);

/* Fused TrafficSignalNetwork.step: apply actions (device int32 [B][A]; ignored
 * when controller == TSC_CTRL_FIXED_TIME), n_ticks engine ticks, then everything
 * tsc_retrieve does.  controller_arg = green time for TSC_CTRL_FIXED_TIME
 * (controllers/controllers.py:26-54). */
int  tsc_env_step(tsc_handle h, const int32_t *actions, int32_t controller, int32_t controller_arg,
                  int32_t n_ticks, const tsc_outputs_t *out, void *stream);

/* What pytsc's rule-based controller `controller` (TSC_CTRL_FIXED_TIME, _GREEDY, _MAX_PRESSURE, _SOTL,
 * _RANDOM) would choose from the current state, without applying it: actions_out device int32 [B][A]
 * pytsc phase indices; scores_out (may be NULL) device int32 [B][A][P]: per phase index the queue
 * (GREEDY, controllers.py:95-114) or pressure (MAX_PRESSURE, :151-176) the reference computes for it,
 * TSC_SCORE_MASKED where the mask forbids it or the signal is on yellow; SOTL: [0] = flow on the current
 * phase, [1] = flow on the next green phase (:222-238), rest 0. */
int  tsc_controller_act(tsc_handle h, int32_t controller, int32_t controller_arg, int32_t *actions_out,
                        int32_t *scores_out, void *stream);

/* Same through HOST buffers: copies `actions_host` in, runs the fused step and
 * copies obs / reward / mask / reward_global back (any may be NULL), then
 * synchronises.  The batch is stepped in chunks of whole grid waves (at most 16
 * chunks, TSC_B200_HOST_CHUNKS) on an internal stream while the observation rows
 * of finished chunks are copied out on another; page-locked caller buffers are
 * used in place.  This is the end-to-end path
 * bench.py times as `e2e`. */
int  tsc_env_step_host(tsc_handle h, const int32_t *actions_host, int32_t controller, int32_t controller_arg,
                       int32_t n_ticks, float *obs_host, float *reward_host, uint8_t *mask_host,
                       float *reward_global_host);

/* Registered end-to-end path.  tsc_host_register takes the caller's HOST result buffers once (ordinary
 * or page-locked memory, any may be NULL): obs [B][A][obs_dim] float32 (lane_features observations only),
 * reward [B][A] float32, mask [B][A][n_actions] uint8, reward_global [B] float32.  It writes the static
 * and padding columns of every observation row (156 of 212 floats per Hangzhou row) once.
 * tsc_env_step_registered then runs the fused step in ONE launch; each replica block stores a compact
 * packet (per incoming lane n_queued / occupancy / mean_speed as the row shows them, per signal the phase
 * index, reward, allowed-action bits; the global reward: < 1 KB per Hangzhou replica instead of 13.6 KB of
 * fp32 rows) straight into page-locked host memory and raises a per-replica flag behind a system fence.
 * Host worker threads (TSC_B200_HOST_THREADS, default min(8, cores / LOCAL_WORLD_SIZE)) follow the flags
 * while the launch is still running and finish the rows in the caller's buffers: only values that changed
 * since the previous step are rewritten.  After the call the registered buffers hold exactly what
 * tsc_env_step + a device-to-host copy of obs / reward / mask / reward_global would (bit for bit).
 * actions_host: host int32 [B][A] (ignored by the in-kernel controllers).  Synchronous.
 * tsc_host_packet_bytes: bytes that cross PCIe device-to-host per step (packets + flags). */
int  tsc_host_register(tsc_handle h, float *obs_host, float *reward_host, uint8_t *mask_host,
                       float *reward_global_host);
int  tsc_host_unregister(tsc_handle h);
int  tsc_env_step_registered(tsc_handle h, const int32_t *actions_host, int32_t controller,
                             int32_t controller_arg, int32_t n_ticks);

/* The same step in two halves, for callers that keep several handles (or their own work) in flight -- e.g. two
 * handles of B/2 replicas each, stepped alternately so that the host policy and row finishing of one half overlap
 * the launch of the other (the double-buffered sampling loop of bench.py's e2e leg).  _begin queues the action
 * copy and the launch and wakes the worker threads, then returns; _wait waits for the workers (they
 * finish all the rows) and the stream, and reports errors.  A page-locked actions_host must stay untouched
 * until _wait returns.  At most one step per handle in flight.  tsc_env_step_registered = _begin + _wait. */
int  tsc_env_step_registered_begin(tsc_handle h, const int32_t *actions_host, int32_t controller,
                                   int32_t controller_arg, int32_t n_ticks);
int  tsc_env_step_registered_wait(tsc_handle h);

/* Worker threads of the registered path for handles registered AFTER this call (0 = the automatic choice,
 * min(8, cores / LOCAL_WORLD_SIZE), or TSC_B200_HOST_THREADS); returns the count that will be used. */
int  tsc_host_threads(tsc_handle h, int32_t n);
int64_t tsc_host_packet_bytes(tsc_handle h);

/* Copy the running vehicles of replica b to host arrays of capacity `cap`
 * (drivable-major, front to back): vehicle id (creation order), drivable,
 * distance, speed.  Returns the vehicle count (may exceed cap) or <0.  Syncs. */
int  tsc_snapshot(tsc_handle h, int32_t replica, int32_t cap, int32_t *vid, int32_t *drivable,
                  double *distance, double *speed, int32_t *blocker_vid, int32_t *enter_ll_time);

/* Load a vehicle snapshot into replica b (test / fixture entry point): n
 * vehicles given drivable-major front to back.  Syncs. */
int  tsc_load_snapshot(tsc_handle h, int32_t replica, int32_t n, const int32_t *vid, const int32_t *drivable,
                       const double *distance, const double *speed, const int32_t *route_pos);

/* MetricsParser.mst (backends/cityflow/metrics.py:202-209; common/utils.py:158-161): the maximum spanning tree
 * (forest, if the signal graph is not connected) of every replica's density map.  density_map: device double
 * [B][A][A] as tsc_retrieve writes it (symmetric; 0 = no edge); out: device double [B][A][A], -w at [min(u,v)][max(u,v)]
 * for every tree edge {u, v} of weight w (the sign scipy's minimum_spanning_tree(-density_map) leaves), 0 elsewhere.
 * Prim's algorithm, one thread block per replica; among equal weights the lower vertex index wins. */
int  tsc_max_spanning_tree(tsc_handle h, const double *density_map, double *out, void *stream);

/* Synchronise and report sticky device-side error flags: returns 0 or the most
 * severe TSC_E* code; *first_bad_replica (may be NULL) gets the replica index. */
int  tsc_check(tsc_handle h, int32_t *first_bad_replica);

/* Per-replica counters (host arrays of length B, any may be NULL).  Syncs. */
int  tsc_counters(tsc_handle h, int32_t *tick, int32_t *n_running, int32_t *n_finished, int32_t *n_slots);

/* Number of kernel launches issued through this handle since creation. */
int64_t tsc_launch_count(tsc_handle h);

/* Debug: per-phase clock64() sums seen by thread 0 of every block since timing was (re)enabled.
 * Copies up to n counters into cycles_out (may be NULL), then enables (zeroing) or disables the
 * instrumentation; returns the number of counters (stage-in, prologue, handleWaiting + look-ahead, retrieve's
 * lane sums, decisions, retrieve's per-signal block, cross phase, list surgery leave / enter, compaction, rest of
 * retrieve, stage-out, then four list-length sums).  Syncs.  Only builds made with
 * -DTSC_PHASE_TIMING carry the instrumentation (tools/phase_timing.py makes one); the shipped library answers
 * TSC_EINVAL to enable = 1. */
int  tsc_debug_timing(tsc_handle h, int32_t enable, uint64_t *cycles_out, int32_t n);

/* Name, bytes of dynamic shared memory, threads per block and grid of the step kernel. */
int  tsc_kernel_info(tsc_handle h, int32_t *smem_bytes, int32_t *threads, int32_t *grid, int32_t *regs);

/* Which variant of the step kernel the handle runs: fixed_capacity = 1 when the kernel was compiled for exactly
 * this handle's vehicle capacity (the bench workloads' capacities: the per-vehicle columns of the working set sit at
 * compile-time offsets; TSC_B200_FIXED_CAPACITY=0 forces the generic build); global_workspace != 0 when the
 * replica's working set does not fit shared memory: 2 = its per-vehicle columns live in a global-memory
 * (L2-resident) workspace, the per-drivable arrays, scratch lists and header in shared memory; 1 = all of it
 * in the workspace (TSC_B200_GMEM_META_SHARED=0, or more drivables than shared memory holds);
 * blocks_per_sm = launch-bounds variant. */
int  tsc_kernel_variant(tsc_handle h, int32_t *fixed_capacity, int32_t *global_workspace, int32_t *blocks_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* TSC_B200_H */
