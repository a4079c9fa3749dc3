// tsc_b200.cu -- B200 (sm_100a) batched traffic-signal-control engine.
//
// One thread block owns one scenario replica for a whole env-step: the
// replica's packed state image is staged HBM -> shared memory, the phase
// controller, `n_ticks` CityFlow-semantics ticks and the Retriever /
// observation / reward / mask reductions all run out of shared memory, and the
// image is written back once.  HBM therefore sees each vehicle once per
// env-step (read + write), not once per tick.  See DESIGN.md for the layout and
// the roofline; include/tsc_b200.h for the ABI and the reference call sites.
//
// Engine semantics: SURVEY.md Appendix A (CityFlow restated).  All kinematics
// are fp64 and compiled with -fmad=false: every operation rounds once, in the
// order written, so that results are bit-identical to the CPU oracle.
#include "tsc_b200.h"

#include <cuda_runtime.h>

#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <sys/prctl.h>
#include <time.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sched.h>

// ----------------------------------------------------------------------------
// Device-side views
// ----------------------------------------------------------------------------
typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned int u32;

// Packed static tables: one 32-byte load per lane-link, one 48-byte load per cross,
// instead of chains of dependent 4-byte loads through seven separate arrays.
struct __align__(16) LLInfo {
    int start_lane, end_lane;
    int cross_off, cross_end;     // range in the cross table, ascending distance along the link
    double length;
    int type;                     // 3 go_straight, 2 turn_left, 1 turn_right
    int sigbit;                   // signal | road-link << 16
};
static_assert(sizeof(LLInfo) == 32, "LLInfo must be 32 bytes");
struct __align__(16) CrossEntry {
    double dist;                  // distance of the cross along this lane-link
    double foe_dist;              // ... along the other lane-link
    double foe_len;               // length of the other lane-link
    double foe_sl_len;            // length of the other link's start lane
    int foe_ll, foe_start_lane, foe_end_lane, foe_type;
};
static_assert(sizeof(CrossEntry) == 48, "CrossEntry must be 48 bytes");

struct DevScn {   // device copies of tsc_scenario_t tables
    const LLInfo *llinfo;
    const CrossEntry *cross;
    int L, K, D, A, N, T, F, horizon, max_raw, P;
    int n_in_total, n_out_total, n_spawn_lanes;
    int spawn_pure;                  // 1: no spawn lane is the end lane of a lane-link (handleWaiting and the look-ahead touch disjoint drivables)
    const double *drv_length, *drv_max_speed;
    const double2 *drv_lm;      // [D] (length, max speed) packed for the per-vehicle pass
    const int *lane_ll_off, *lane_ll, *lane_spawn_off, *lane_spawn_vid, *spawn_lane;
    const short *lane_spawn_idx;     // [L] index into the spawn-lane list, -1 for lanes nothing spawns on
    const int4 *lane_sib;            // [L] {n, d0, d1, d2}: the (up to three) lane-links leaving the lane as drivable indices; n = -1: more, use lane_ll
    const int *ll_start_lane, *ll_end_lane, *ll_signal, *ll_roadlink, *ll_type, *ll_cross_off;
    const double *xr_dist, *xr_foe_dist;
    const int *xr_foe_ll;
    const u32 *sig_phase_mask;
    const int *sig_n_raw;            // [A] raw light phases per signal
    const int *route_seq, *veh_tick, *veh_seq_start, *veh_tmpl, *veh_priority;
    const int4 *spawn_rec;           // [N] in lane_spawn_vid order: {vehicle, creation tick, route cursor, second drivable}
    const double *tmpl;
    const int *created_cnt;          // [F][horizon+2] vehicles of flow set f created before tick t
    const long long *created_enter;  // [F][horizon+2] sum of their creation ticks
    // host packet of the registered end-to-end path (tsc_env_step_registered)
    const u32 *pk_lane;              // [n_in_total] incoming lanes in observation-row order: lane | truncate << 31
    int pk_mode;                     // 1: one u32 per lane (n_queued | occupancy << 8 | mean_speed << 16, integers); 0: three floats
    int pk_o_phase, pk_o_reward, pk_o_mask, pk_o_rg, pk_bytes;      // byte offsets inside a replica's packet; its size (multiple of 16)
    // pytsc tables
    const double *lane_pytsc_length, *lane_feat, *lane_cells;
    const int *sig_in_off, *sig_in_lane, *sig_out_off, *sig_out_lane;
    const int *in_sig;               // [n_in_total] the signal an incoming-lane entry belongs to
    const int *sig_n_phases, *sig_phase_raw, *sig_min_time, *sig_max_time;
    const u8 *sig_phase_green;
    const int *nbr_off, *nbr_idx;
    const double *nbr_weight;
    const int *ctl_off, *ctl_in_lane, *ctl_out_lane;   // rule-based controllers: lanes served by every pytsc phase
    const int *dm_off, *dm_lane;                       // density map: lanes leading from signal i to signal j
    const double *dm_adjacency;
    const u32 *obs_code;      // [A * state_dim] lane-feature row recipe: 0 = static, else kind | truncate << 3 | argument << 4
    const float *obs_static;  // [A * state_dim] the static values (lane features as the reference stores them, -1 padding)
    int reward_type, obs_type, action_space, round_robin, visibility, yellow_time;
    int obs_dim, state_dim, n_actions, reference_exact, max_lanes_per_signal, max_obs_phases;
    double v_size, flick, interval;
};

struct RepHeader {   // 64 bytes, first thing in every replica image
    int n_slots;          // slots in use: live vehicles and holes left by finished ones (compacted away now and then)
    int tick;             // engine step counter
    int n_running;
    int n_finished;
    long long cum_tt;     // sum over finished vehicles of (finish tick - creation tick)
    long long fin_enter;  // sum of creation ticks of finished vehicles
    u32 err;              // sticky error bits
    int n_ent;            // scratch: vehicles changing drivable (or finishing) this tick  } even ticks; odd ticks count in
    int n_x;              // scratch: vehicles deferred to the cross phase this tick       } n_h / n_a (tick_counters)
    int n_h, n_a;         // scratch: the same two counters of odd ticks
    int n_new;            // scratch: n_slots after this tick's spawns
    int flow_set;         // which of the scenario's flow sets this replica runs (set at reset)
    int pad;
};
static_assert(sizeof(RepHeader) == 64, "RepHeader must be 64 bytes");

#define ERR_OVERFLOW 1u
#define ERR_ORDER 2u
#define ERR_ENT_OVERFLOW 4u
#define ERR_BAD_PHASE 8u      // tsc_set_phase / actions named a phase the signal does not have

#define NONE16 0xFFFFu        // "no vehicle" in the per-drivable lists
#define PJ_MOVER 0x80u        // pj[] bit: the vehicle leaves its drivable this tick (set by its decision, cleared by the list surgery)
#define HOLE_MAX 32           // finished vehicles leave holes; the slots are compacted once this many have piled up
#define HOLE_MIN_ROUND 6      // ... or this many, when squeezing them out saves a round of the per-vehicle passes

struct Layout {
    int Vcap;              // vehicle slots (running vehicles + holes)
    int Vlay;              // slots the columns are laid out for (>= Vcap: the retrieve scratch lives in a pos/spd pair)
    int ent_cap;           // vehicles that may change drivable in one tick
    int wl_cap;            // entries of a warp's private head-vehicle list
    int async_stage;       // staging of the image columns (TSC_B200_ASYNC_STAGE): 2 bulk asynchronous copies + mbarrier, 1 cp.async, 0 plain
    int prefetch_next;     // 1: L2 prefetch of the block's next replica image during the step
    int warp_surgery;      // 1: up to 32 movers do both halves of the list surgery inside the first warp (TSC_B200_WARP_SURGERY)
    int meta_shared;       // global-memory variant: everything from o_cnt on (per-drivable arrays, scratch, header) in shared memory
    // Per-vehicle columns first: their offsets depend on the capacity and the block size alone (v_offsets), so a kernel
    // built for one capacity has them as compile-time constants.  The hot columns have identical byte offsets in the HBM
    // image and in the working set.
    int o_pos, o_spd, o_rpos, o_vid, o_lead, o_foll, o_drv, o_pj, o_blk;
    // image only, behind the hot columns: the per-drivable / per-signal block (o_imeta, meta_bytes), then the cold column
    // (enterLaneLinkTime), which is worked on in place in the image
    int o_imeta, meta_bytes, o_ellt, img_bytes;
    // working set only, behind the hot columns: more per-vehicle columns, then the per-drivable / per-signal block (a
    // straight copy of the image's), then scratch
    int o_dn, o_npos, o_nspd, o_nblk, o_xlist;
    int o_cnt, o_head, o_tail, o_wq, o_sraw, o_scur, o_schg, o_stop, o_meta_end;
    int o_leave, o_ent, o_fresh, o_mvslot, o_mvto, o_mvq, o_mvpj, o_scan, o_avail, o_tmpl, o_spawn, smem_bytes;
};

// Offsets of the per-vehicle columns for V slots (columns laid out for Vlay >= V slots) and n_warps warps per block.
struct VOff { int pos, spd, rpos, vid, lead, foll, drv, pj, blk, hot_end, dn, npos, nspd, nblk, xlist, end, wl_cap; };
__host__ __device__ constexpr int v_align16(int x) { return (x + 15) & ~15; }
__host__ __device__ constexpr VOff v_offsets(int V, int Vlay, int n_warps) {
    VOff r{};
    int o = (int) sizeof(RepHeader);
    r.pos = o; o = o + 8 * Vlay;                // pos | spd contiguous
    r.spd = o; o = v_align16(o + 8 * Vlay);
    r.rpos = o; o = v_align16(o + 4 * V);
    r.vid = o; o = v_align16(o + 4 * V);
    r.lead = o; o = v_align16(o + 2 * V);
    r.foll = o; o = v_align16(o + 2 * V);
    r.drv = o; o = v_align16(o + 2 * V);        // HBM: u16 drivable; working set: the list of the cross phase
    r.pj = o; o = v_align16(o + V);
    r.blk = o; o = v_align16(o + 2 * V);
    r.hot_end = o;
    r.dn = o; o = v_align16(o + 4 * V);
    r.npos = o; o = o + 8 * Vlay;               // npos | nspd contiguous
    r.nspd = o; o = v_align16(o + 8 * Vlay);
    r.nblk = o; o = v_align16(o + 2 * V);
    r.wl_cap = (((V + 31) / 32 + n_warps - 1) / n_warps) * 32;      // a warp owns every n_warps-th chunk of 32 slots
    r.xlist = o; o = v_align16(o + 2 * (V > n_warps * r.wl_cap ? V : n_warps * r.wl_cap));
    r.end = o;
    return r;
}

struct StepArgs {
    int b0;                // first replica of this launch
    int B;                 // one past the last replica of this launch
    int n_ticks;
    int apply_actions;     // 0 none, 1 external actions, 2 fixed-time controller, 3 external phase indices,
                           // 4 greedy, 5 max pressure, 6 SOTL, 7 random (rule-based controllers on the device)
    int controller_arg;
    int decide_only;       // 1: report what the controller would choose, leave the programs alone
    int *ctl_actions;      // [B][A] out: the controller's phase indices, or NULL
    int *ctl_scores;       // [B][A][P] out: per-phase scores behind the decision, or NULL
    int do_retrieve;
    int set_raw_phase;     // 1: raw_phase input given
    int init_program;      // >=0: TSProgram.set_initial_phase(index)
    const int *actions;    // [B][A]
    const int *raw_phase;  // [B][A]
    unsigned long long *phase_cycles;   // debug: per-phase clock64 sums (thread 0 of every block), or NULL
    unsigned char *workspace;           // GMEM variant: one working set of Y.smem_bytes (256-byte aligned stride) per block
    unsigned char *pk;                  // registered host path: packets [B][pk_bytes] in mapped page-locked host memory, or NULL
    u32 *pk_flags;                      // [B] raised to pk_seq (behind a system fence) once replica b's packet is complete
    u32 pk_seq;
    tsc_outputs_t out;
};

// ----------------------------------------------------------------------------
// Small device helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ double min2(double x, double y) { return x < y ? x : y; }
__device__ __forceinline__ double max2(double x, double y) { return x > y ? x : y; }

// (int) of a double as the reference's x86 build performs it: out-of-range and
// NaN give INT_MIN (cvttsd2si), in-range truncates toward zero.
__device__ __forceinline__ int trunc_int_x86(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT_MIN;
    return (int) x;
}

#define SMEM_TEMPLATES 4      // vehicle templates cached in shared memory (more: read from global)

struct Ctx {
    RepHeader *h;
    unsigned char *vb;    // base of the per-vehicle columns (shared memory, or the block's global-memory workspace)
    // per drivable: vehicle count and the ends of its list (front = head, back = tail)
    u16 *cnt, *head, *tail, *wq;
    u8 *leave, *ent;      // per drivable: vehicles leaving / entering this tick (bytes, counted four to an atomic word)
    // per vehicle slot.  Slots are STABLE: a vehicle keeps its slot from spawn to finish; order on a drivable is
    // the doubly linked list lead (vehicle ahead) / foll (vehicle behind).
    u16 *lead, *foll;
    u16 *xlist;           // the warps' private head-vehicle lists (wl_cap entries each); the slot map of a compaction
    u16 *clist;           // the block's list of vehicles that go through the cross phase this tick
    u16 *mv_slot, *mv_to; // this tick's movers: slot, new drivable (NONE16 = route ends)
    int *mv_q;            // ... new route cursor
    u8 *mv_pj;            // ... skipped a whole drivable
    u32 *dn;              // per vehicle: drivable | next drivable << 16 (0xFFFF = route ends)
    u32 *avail;           // bit per lane-link: its road-link is green in the signal's current phase
    const double *tmpl;   // vehicle templates (shared-memory copy when they fit)
    u8 *sraw, *scur, *schg, *pj, *fresh;
    int *stop, *rpos, *vid, *ellt, *scan;
    int *sp_lane;                      // [S] the spawn lanes
    int *sp_rec;                       // [S] int4: head of every spawn lane's waiting buffer (vehicle, creation tick, route cursor, second drivable)
    int *sp_base;                      // [2S] this replica's flow set: first / one-past-last spawn record of every spawn lane
    const int *lso;                    // this replica's flow set: row of lane_spawn_off
    // kinematics ping-pong: every tick reads pos / spd / blk and writes npos / nspd / nblk, then they swap
    short *blk, *nblk;
    double *pos, *spd, *npos, *nspd;
    int par;              // which half of the ping-pong holds the current state (0: the one that mirrors the image)
    int tick;
#ifdef TSC_PHASE_TIMING
    unsigned long long *pt;   // debug phase timing (NULL = off)
    long long pt_last;
#endif
};

// phase ids of the debug timing
enum { PT_STAGE_IN = 0, PT_PROLOGUE, PT_SPAWN, PT_PHASE1A, PT_PHASE1, PT_PHASE1C, PT_PHASE2, PT_LEAVE, PT_ENTER, PT_COMPACT, PT_RETRIEVE,
       PT_STAGE_OUT, PT_NH, PT_NA, PT_NX, PT_NENT, PT_N };
// (compiled in only with -DTSC_PHASE_TIMING, tools/phase_timing.py's build: in the shipped kernel the two words would
// live on the stack and be reloaded beside every barrier)
__device__ __forceinline__ void pt_mark(Ctx &c, int k) {
#ifdef TSC_PHASE_TIMING
    if (c.pt && threadIdx.x == 0) { long long t = clock64(); atomicAdd(c.pt + k, (unsigned long long) (t - c.pt_last)); c.pt_last = t; }
#endif
}

// The kinematics ping-pong: which buffers are "current" and which "next" follows from the parity alone, so a tick ends
// by flipping one integer (swapping six pointers that live on the stack cost a dependent local load + store each).
__device__ __forceinline__ void set_pingpong(const Layout &Y, Ctx &c) {
    unsigned char *const smem = c.vb;
    const bool p = c.par != 0;
    c.pos = (double *) (smem + (p ? Y.o_npos : Y.o_pos)); c.npos = (double *) (smem + (p ? Y.o_pos : Y.o_npos));
    c.spd = (double *) (smem + (p ? Y.o_nspd : Y.o_spd)); c.nspd = (double *) (smem + (p ? Y.o_spd : Y.o_nspd));
    c.blk = (short *) (smem + (p ? Y.o_nblk : Y.o_blk)); c.nblk = (short *) (smem + (p ? Y.o_blk : Y.o_nblk));
}

// Append to a per-tick work list from whichever lanes of the warp are here together: one atomic per warp
// instead of one per lane (the counters live in shared memory; a hundred serialised atomics per phase showed up
// as thousands of cycles per tick).  Returns the caller's position in the list.
__device__ __forceinline__ int warp_append(int *counter) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// Device-side vehicle template row: the ABI's TSC_T_STRIDE doubles followed by constants derived on
// the host with the same IEEE operations the CityFlow formulas would repeat for every vehicle.
enum { TD_A = TSC_T_STRIDE,      // 0.5 / maxNegAcc            ("a" of the no-collision quadratic)
       TD_HALF_OVER_A,           // 0.5 / a
       TD_HEADWAY_DEN,           // headwayTime + interval / 2
       TD_BRAKE0,                // stop_before_speed's braking distance for a standing vehicle
       TD_STRIDE };
// The engine runs at CityFlow's interval of 1.0 s only (tsc_create rejects anything else), so
// x * DT and x / DT are exact identities and fold away.
#define DT 1.0

// ONE_T: the scenario has a single vehicle template (every shipped flow file): the row is the
// shared-memory copy at a fixed address, and the vehicle-id loads that only selected it disappear.
template <bool ONE_T>
__device__ __forceinline__ const double *tmpl_of(const DevScn &S, const Ctx &c, int vid) {
    if (ONE_T) return c.tmpl;
    return c.tmpl + TD_STRIDE * __ldg(S.veh_tmpl + vid);
}

// ---- A.4 car following -------------------------------------------------------
// a = 0.5 / dF and 0.5 / a come from the follower's template row.
// x / y for y > 0 finite, x >= 0.  The division's exponent-range check sends zero numerators to its
// out-of-line slow path (~60 instructions), and standing vehicles / empty lanes make them the common
// case.  A plain `x == 0 ? x : x / y` is if-converted by the compiler (the division, slow path
// included, runs anyway and a select picks the result), so the zero is replaced by 1.0 before the
// division and put back after it: the quotient of a zero numerator is that zero, bit for bit.
__device__ __forceinline__ double div_pos(double x, double y) {
    const bool zero = x == 0.0;
#ifndef TSC_DIV_POS_NO_BARRIER
    // The empty asm makes the substituted value opaque: without it the optimiser proves the substitution
    // dead and divides x itself, and the slow path still runs for every zero numerator (SASS of an earlier
    // build: 1.03 M calls = 14 % of the executed instructions per launch).  0.913 -> 0.848 ms.
    double xs = zero ? 1.0 : x;
    asm volatile("" : "+d"(xs));
    return zero ? x : xs / y;
#else
    const double q = (zero ? 1.0 : x) / y;
    return zero ? x : q;
#endif
}
// the three other places where standing vehicles divide a zero (0.5 v^2 / maxNegAcc)
#ifndef TSC_DIV_POS_NO_BARRIER
#define DIV_POS_HOT(x, y) div_pos((x), (y))
#else
#define DIV_POS_HOT(x, y) ((x) / (y))
#endif

// dL > 0 (a deceleration the caller knows to be positive: a template's maxNegAcc, or v - vL > 0)
__device__ __forceinline__ double no_collision_speed(double vL, double dL, double vF, double a, double half_over_a, double gap,
                                                     double target) {
    double c = vF * DT / 2 + target - div_pos(0.5 * vL * vL, dL) - gap;
    double b = 0.5 * DT;
    if (b * b < 4 * a * c) return -100;
    double v1 = half_over_a * (sqrt(b * b - 4 * a * c) - b);
    double v2 = 2 * vL - dL * DT + 2 * (gap - target) / DT;
    return min2(v1, v2);
}

__device__ __forceinline__ double car_follow_speed(const double *T, double v, double gap, double vL, double leaderMaxNegAcc) {
    double s = no_collision_speed(vL, leaderMaxNegAcc, v, T[TD_A], T[TD_HALF_OVER_A], gap, 0);
    double assumeDecel = 0, s2;
    if (v > vL) {
        assumeDecel = v - vL;
        s2 = no_collision_speed(vL, assumeDecel, v, T[TD_A], T[TD_HALF_OVER_A], gap, T[TSC_T_MIN_GAP]);
    } else {
        // assumeDecel == 0: CityFlow's formula divides by it.  0.5 vL^2 / 0 is +inf (or NaN for vL^2 == 0), so c is
        // -inf (NaN), the discriminant test is false, v1 = half_over_a * (sqrt(+inf | NaN) - b) is +inf (NaN) because
        // a, half_over_a > 0 (checked by tsc_create), and min2(v1, v2) = (v1 < v2 ? v1 : v2) returns v2 either way:
        // the same bits without the division, the square root and their slow paths.
        s2 = 2 * vL - 0.0 * DT + 2 * (gap - T[TSC_T_MIN_GAP]) / DT;
    }
    s = min2(s, s2);
    s = min2(s, (gap + (vL + assumeDecel / 2) * DT - v * DT / 2) / T[TD_HEADWAY_DEN]);
    return s;
}

__device__ double stop_before_speed(const double *T, double v, double distance) {
#ifndef TSC_NO_STANDING_SHORTCUT
    // A standing vehicle (most callers: the queue in front of a red light) not exactly on the line: brake is the
    // template constant TD_BRAKE0 (the same operations on v = 0, done once on the host); if it does not fit, take =
    // 2 distance / 1e-8 >= 200 and the answer is 0 - 0 / n = +0 whatever n (even cvttsd2si's overflow value): three
    // divisions, mostly on their slow paths, answered by a compare.
    if (v == 0.0 && distance > 1e-6) return T[TD_BRAKE0] < distance ? T[TSC_T_USUAL_POS_ACC] : 0.0;
#endif
    double nxt = v + T[TSC_T_USUAL_POS_ACC] * DT;
    double brake = (v + nxt) * DT / 2 + (nxt * nxt / T[TSC_T_USUAL_NEG_ACC] / 2);
    if (brake < distance) return v + T[TSC_T_USUAL_POS_ACC] * DT;
    double take = 2 * distance / (v + 1e-8) / DT;
    if (take >= 1) return v - v / trunc_int_x86(take);
    return v - v / take;
}

__device__ __forceinline__ bool can_yield(const double *T, double v, double dist) {
    double minBrake = DIV_POS_HOT(0.5 * v * v, T[TSC_T_MAX_NEG_ACC]);
    return (dist > 0 && minBrake < dist - T[TSC_T_YIELD_DIST]) || (dist < 0 && dist + T[TSC_T_LEN] < 0);
}

__device__ int reach_steps(const double *T, double v, double distance, bool turn) {
    const double dt = DT;
    double target = turn ? T[TSC_T_TURN_SPEED] : T[TSC_T_MAX_SPEED];
    double acc = T[TSC_T_USUAL_POS_ACC];
    if (distance <= 0) return -1;
    if (v > target) return trunc_int_x86(ceil(distance / v));
    double dUntil = 0;
    if (!(target <= v)) {
        int s1 = trunc_int_x86(floor((target - v) / acc / dt));
        double s1speed = v + s1 * acc / dt;
        double s1dis = (v + s1speed) * (s1 * dt) / 2;
        dUntil = s1dis + (s1speed < target ? ((s1speed + target) * dt / 2) : 0);
    }
    if (dUntil > distance) return trunc_int_x86(ceil((sqrt(v * v + 2 * acc * distance) - v) / acc / dt));
    return trunc_int_x86(ceil((target - v) / acc / dt)) + trunc_int_x86(ceil((distance - dUntil) / target / dt));
}

__device__ __forceinline__ bool ll_available(const Ctx &c, int ll) { return (c.avail[ll >> 5] >> (ll & 31)) & 1u; }

// Which vehicle does the lane-link described by `X` (the foe side of a cross)
// announce at that cross (CityFlow notifyCross; closed form of the sequential
// scan)?  Returns the slot or -1; *d2 = its signed distance to
// the cross.  Conditions are ordered so that shared memory decides first and
// the route table (global) is read only when everything else already holds.
template <bool ONE_T>
__device__ int cross_claimant(const DevScn &S, const Ctx &c, const CrossEntry &X, double *d2) {
    const int f = X.foe_ll, fl = S.L + f;
    const double dc = X.foe_dist;
    if (c.cnt[X.foe_end_lane] > 0) {   // the vehicle that has just moved onto the end lane
        const int t = c.tail[X.foe_end_lane];
        double crossDistance = X.foe_len - dc;
        double vehDistance = c.pos[t] - tmpl_of<ONE_T>(S, c, c.vid[t])[TSC_T_LEN];
        if (crossDistance + vehDistance < 0.0 && !(c.pj[t] & 1) && __ldg(S.route_seq + c.rpos[t] - 1) == fl) {
            *d2 = -(c.pos[t] + crossDistance);
            return t;
        }
    }
    int n = c.cnt[fl];
    for (int v = n > 0 ? (int) c.head[fl] : (int) NONE16; n > 0 && v != (int) NONE16; --n, v = c.foll[v]) {   // vehicles on the link, front to back
        double vd = c.pos[v];
        if (vd > dc) {
            if (vd - dc - tmpl_of<ONE_T>(S, c, c.vid[v])[TSC_T_LEN] <= 0.0) { *d2 = dc - vd; return v; }
        } else { *d2 = dc - vd; return v; }
    }
    if (c.cnt[X.foe_start_lane] > 0 && ll_available(c, f)) {   // first vehicle of the incoming lane, heading here on green
        int hd = c.head[X.foe_start_lane];
        if ((int) (c.dn[hd] >> 16) == fl) {
            *d2 = (X.foe_sl_len - c.pos[hd]) + dc;
            return hd;
        }
    }
    return -1;
}

// Cross::canPass (A.5) for vehicle `me` on / approaching a lane-link of type t1,
// at the cross `X`.  *foe_out = announced vehicle on the other link.
template <bool ONE_T>
__device__ bool can_pass(const DevScn &S, const Ctx &c, int me, const double *T, int t1, const CrossEntry &X, double dts,
                         int *foe_out) {
    double d2;
    int foe = cross_claimant<ONE_T>(S, c, X, &d2);
    *foe_out = foe;
    if (foe < 0) return true;
    int t2 = X.foe_type;
    double d1 = X.dist - dts;
    double v = c.spd[me];
    if (!can_yield(T, v, d1)) return true;
    const double *TF = tmpl_of<ONE_T>(S, c, c.vid[foe]);
    double vf = c.spd[foe];
    int yield = 0;
    if (!can_yield(TF, vf, d2)) yield = 1;
    if (yield == 0) {
        if (t1 > t2) yield = -1;
        else if (t1 < t2) {
            if (d2 > 0) {
                int fs = reach_steps(TF, vf, d2, t2 != 3);
                int ms = reach_steps(T, v, d1, t1 != 3);
                if (fs > ms) yield = -1;
            } else if (d2 + TF[TSC_T_LEN] < 0) yield = -1;
            if (yield == 0) yield = 1;
        } else {
            if (d2 > 0) {
                int fs = reach_steps(TF, vf, d2, t2 != 3);
                int ms = reach_steps(T, v, d1, t1 != 3);
                if (fs > ms) yield = -1;
                else if (fs < ms) yield = 1;
                else {
                    int e1 = c.ellt[me], e2 = c.ellt[foe];
                    if (e1 == e2) {
                        if (d1 == d2) yield = __ldg(S.veh_priority + c.vid[me]) > __ldg(S.veh_priority + c.vid[foe]) ? -1 : 1;
                        else yield = d1 < d2 ? -1 : 1;
                    } else yield = e1 < e2 ? -1 : 1;
                }
            } else yield = d2 + TF[TSC_T_LEN] < 0 ? -1 : 1;
        }
    }
    if (yield == 1) {   // deadlock: the foe's blocker chain loops back
        int fast = foe, slow = foe;
        while (fast >= 0 && c.blk[fast] >= 0) {
            slow = c.blk[slow];
            fast = c.blk[c.blk[fast]];
            if (slow == fast) { yield = -1; break; }
        }
    }
    return yield == -1;
}

// The two per-tick list counters alternate between two pairs of header words by tick parity: a tick's first phase zeroes
// the pair of the NEXT tick (idle for the whole of this one), so no reset ever shares a phase with a reader or a writer.
// (the header's tick is advanced inside the list surgery: only the decisions may read the parity from there)
__device__ __forceinline__ int *mover_counter(const Ctx &c, int tick) { return (tick & 1) ? &c.h->n_h : &c.h->n_ent; }
__device__ __forceinline__ int *cross_counter(const Ctx &c, int tick) { return (tick & 1) ? &c.h->n_a : &c.h->n_x; }

// Commit one vehicle's decision into the tick's next-state buffers: clamp the speed, advance along the
// route, and register the move if it leaves its drivable (A.4).
__device__ __forceinline__ void finish_vehicle(const DevScn &S, const Layout &Y, Ctx &c, int i, const double *T, int d, int rp,
                                               double x, double v, double dlen, double ns, int blocker) {
    const double dt = DT;
    ns = max2(ns, v - T[TSC_T_MAX_NEG_ACC] * dt);
    double delta;
    if (ns < 0) { delta = DIV_POS_HOT(0.5 * v * v, T[TSC_T_MAX_NEG_ACC]); ns = 0; }
    else delta = (v + ns) * dt / 2;
    double nx = delta + x;
    int q = rp, dd = d, hops = 0;
    bool end = false;
    double cl = dlen;
    while (nx > cl) {   // walk forward through the drivables
        nx -= cl;
        int nxt = __ldg(S.route_seq + q + 1);
        if (nxt < 0) { end = true; break; }
        ++q; dd = nxt; ++hops;
        cl = __ldg(S.drv_length + dd);
    }
    c.npos[i] = nx; c.nspd[i] = ns; c.nblk[i] = (short) blocker;
    if (hops > 0 || end) {           // leaves its drivable: the list surgery at the end of the tick moves it
        c.pj[i] |= PJ_MOVER;
        atomicAdd((unsigned *) (c.leave + (d & ~3)), 1u << (8 * (d & 3)));
        if (!end) atomicAdd((unsigned *) (c.ent + (dd & ~3)), 1u << (8 * (dd & 3)));
        const int k = atomicAdd(mover_counter(c, c.h->tick), 1);
        if (k < Y.ent_cap) {
            c.mv_slot[k] = (u16) i; c.mv_to[k] = end ? (u16) NONE16 : (u16) dd; c.mv_q[k] = q; c.mv_pj[k] = hops > 1 ? 1 : 0;
        } else atomicOr(&c.h->err, ERR_ENT_OVERFLOW);
    }
}

// ---- stable compaction of the vehicle slots (holes left by finished vehicles squeezed out) ----
// Rare: runs at the start of a tick once HOLE_MAX holes have piled up, or when the spawns of the tick could
// run out of slots.  Live slots keep their relative order; every slot reference (list links, blockers, list
// ends) is renumbered.  Rounds of NT slots: a round's sources are read into registers before any of its
// destinations (all at or below the sources) is written.
template <int NT>
__device__ void compact_slots(const DevScn &S, const Layout &Y, Ctx &c) {
    const int tid = threadIdx.x;
    const int n = c.h->n_slots;
    u16 *map = c.xlist;          // dead between ticks
    // exclusive scan of the live flags: thread t owns the contiguous chunk [t * per, (t + 1) * per)
    const int per = (n + NT - 1) / NT;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int live = 0;
    for (int i = lo; i < hi; ++i) live += c.vid[i] >= 0;
    const int lane = tid & 31, w = tid >> 5;
    int incl = live;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) c.scan[w] = incl;
    __syncthreads();
    if (w == 0) {
        int ws = lane < NT / 32 ? c.scan[lane] : 0;
        int wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < NT / 32) c.scan[lane] = wi - ws;
        if (lane == NT / 32 - 1) c.scan[32] = wi;
    }
    __syncthreads();
    int run = c.scan[w] + incl - live;
    for (int i = lo; i < hi; ++i) {
        const bool lv = c.vid[i] >= 0;
        map[i] = lv ? (u16) run : (u16) NONE16;
        run += lv;
    }
    __syncthreads();
    const int n_live = c.scan[32];
    for (int d = tid; d < S.D; d += NT) {
        if (c.cnt[d] > 0) { c.head[d] = map[c.head[d]]; c.tail[d] = map[c.tail[d]]; }
    }
    for (int i0 = 0; i0 < n; i0 += NT) {
        const int i = i0 + tid;
        const int dst = i < n ? (int) map[i] : (int) NONE16;
        double p = 0, s = 0;
        int rp = 0, vd = 0, el = 0;
        u32 dnv = 0;
        u16 ld = NONE16, fo = NONE16;
        short bk = -1;
        u8 pjv = 0;
        if (dst != (int) NONE16) {
            p = c.pos[i]; s = c.spd[i]; rp = c.rpos[i]; vd = c.vid[i]; dnv = c.dn[i]; pjv = c.pj[i];
            el = c.ellt[i];
            ld = c.lead[i]; fo = c.foll[i];
            if (ld != NONE16) ld = map[ld];
            if (fo != NONE16) fo = map[fo];
            int b = c.blk[i];
            if (b >= 0) { const u16 nb = map[b]; b = nb == NONE16 ? -1 : (int) nb; }
            bk = (short) b;
        }
        __syncthreads();
        if (dst != (int) NONE16) {
            c.pos[dst] = p; c.spd[dst] = s; c.rpos[dst] = rp; c.vid[dst] = vd; c.dn[dst] = dnv; c.pj[dst] = pjv;
            c.ellt[dst] = el;
            c.lead[dst] = ld; c.foll[dst] = fo; c.blk[dst] = bk;
        }
    }
    __syncthreads();
    for (int i = n_live + tid; i < n; i += NT) { c.vid[i] = -1; c.blk[i] = -1; c.nblk[i] = -1; }
    if (tid == 0) c.h->n_slots = n_live;
    __syncthreads();
}

// ----------------------------------------------------------------------------
// One engine tick for the replica held in the working set (A.2)
// ----------------------------------------------------------------------------
// Leader and gap of a head vehicle (no vehicle ahead on its drivable) as of the end of the previous tick (A.7): the
// last vehicle of the drivables ahead on its route, as far as it looks; vehicles that left a waiting buffer this tick
// are not yet visible.
template <bool ONE_T>
__device__ __forceinline__ void head_look_ahead(const DevScn &S, const Ctx &c, int i, int *leader_out, double *gap_out) {
    const int L = S.L;
    const double *T = tmpl_of<ONE_T>(S, c, c.vid[i]);
    const u32 dnv = c.dn[i];
    const int d = dnv & 0xFFFF;
    const int nd1 = (dnv >> 16) == 0xFFFFu ? -1 : (int) (dnv >> 16);
    const int rp = c.rpos[i];
    int leader = -1;
    double gap = 0.0;
    double dist = __ldg(S.drv_length + d) - c.pos[i];
    const double horizon = T[TSC_T_APPROACH_DIST];   // maxSpeed^2 / usualNegAcc / 2 + maxSpeed * interval * 2
    for (int j = 1;; ++j) {
        int nd = j == 1 ? nd1 : __ldg(S.route_seq + rp + j);
        if (nd < 0) break;
        if (nd >= L) {
            const int sl = (j == 1 && d < L) ? d : __ldg(&S.llinfo[nd - L].start_lane);   // the link after lane d starts at d
            // all lane-links leaving that lane, in roadnet order: one packed load for up to three of them
            const int4 sib = __ldg(S.lane_sib + sl);
            auto consider = [&](int dl) {
                if (c.cnt[dl] > 0) {
                    int cand = c.tail[dl];
                    double cg = dist + c.pos[cand] - tmpl_of<ONE_T>(S, c, c.vid[cand])[TSC_T_LEN];
                    if (leader < 0 || cg < gap) { leader = cand; gap = cg; }
                }
            };
            if (sib.x >= 0) {
                if (sib.x > 0) consider(sib.y);
                if (sib.x > 1) consider(sib.z);
                if (sib.x > 2) consider(sib.w);
            } else {
                int e0 = __ldg(S.lane_ll_off + sl), e1 = __ldg(S.lane_ll_off + sl + 1);
                for (int q = e0; q < e1; ++q) consider(L + __ldg(S.lane_ll + q));
            }
            if (leader >= 0) break;
        } else {
            const int n = c.cnt[nd] - c.fresh[nd];
            if (n > 0) {      // the lane's last vehicle, not counting one that left the waiting buffer this tick
                int t = c.tail[nd];
                if (c.fresh[nd]) t = c.lead[t];
                leader = t;
                gap = dist + c.pos[leader] - tmpl_of<ONE_T>(S, c, c.vid[leader])[TSC_T_LEN];
                break;
            }
        }
        dist += __ldg(S.drv_length + nd);
        if (dist > horizon) break;
    }
    *leader_out = leader; *gap_out = gap;
}

#define TICK_SYNC_ERR(x) (__syncthreads_or(x) != 0)
// `frozen`: the replica carries a sticky error (the same answer in every thread: it comes out of a barrier).  Returns
// that answer as of the end of the tick.
template <int NT, bool ONE_T>
__device__ bool engine_tick(const DevScn &S, const Layout &Y, Ctx &c, bool frozen) {
    const int tid = threadIdx.x;
    const int tick = c.h->tick;
    const double dt = DT;
    const int L = S.L;
    if (frozen) {   // a replica that overflowed or lost its order is frozen: its result is reported invalid by tsc_check
        __syncthreads();
        if (tid == 0) c.h->tick = tick + 1;
        __syncthreads();
        return true;
    }
    {   // holes left by finished vehicles are squeezed out now and then (uniform decision: every thread reads the same header)
        const int ns = c.h->n_slots, holes = ns - c.h->n_running;
        // ... and as soon as a few of them cost every warp-owned pass over the slots an extra round (a round of NT slots
        // takes as long with one vehicle in it as with NT)
#ifndef TSC_NO_ROUND_COMPACTION
        const bool extra_round = holes >= HOLE_MIN_ROUND && (ns + NT - 1) / NT > (ns - holes + NT - 1) / NT;
#else
        const bool extra_round = false;
#endif
        if (holes >= HOLE_MAX || extra_round || (holes > 0 && ns + S.n_spawn_lanes > Y.Vcap)) {
            compact_slots<NT>(S, Y, c);
            pt_mark(c, PT_COMPACT);
        }
    }
    const int n_old = c.h->n_slots;

    // ---- handleWaiting: at most one vehicle per lane leaves its waiting buffer.  The head of every
    //      buffer (vehicle, creation tick, route cursor, second drivable) is cached in the working set, so a
    //      tick without an arrival costs one compare per spawn lane and an arrival needs no table look-up
    //      before the vehicle is in place; the cache is refilled with ONE 16-byte load whose latency the rest
    //      of the work hides.  The first warp serves the spawn lanes 32 at a time: a new vehicle's slot is
    //      n_slots + its rank among the lanes that spawn (ballot), so slot numbers do not depend on timing. ----
    // fused (-DTSC_FUSED_SPAWN; measured slower, 0.703 vs 0.690 ms, so off): when handleWaiting touches nothing the decisions
    // of the other vehicles read (spawn_pure, see below), the first warp also decides for the vehicles it has just let in, and
    // no block-wide barrier separates this phase from the decisions
#ifdef TSC_FUSED_SPAWN
    const bool fused = S.spawn_pure != 0;
#else
    const bool fused = false;
#endif
    int n_new_w0 = n_old;      // (first warp) slots in use after this tick's spawns
    if (tid < 32) {
        int base = n_old, spawned = 0;
        for (int s0 = 0; s0 < S.n_spawn_lanes; s0 += 32) {
            const int s = s0 + tid;
            bool want = false;
            int l = 0, n = 0;
            int4 rec = make_int4(0, INT_MAX, 0, 0);      // vehicle, creation tick, route cursor, second drivable
            if (s < S.n_spawn_lanes) {
                l = c.sp_lane[s];
                rec = ((const int4 *) c.sp_rec)[s];
                if (rec.y <= tick) {
                    n = c.cnt[l];
                    want = true;
                    if (n > 0) {
                        const int t = c.tail[l];
                        want = c.pos[t] > tmpl_of<ONE_T>(S, c, c.vid[t])[TSC_T_LEN] + tmpl_of<ONE_T>(S, c, rec.x)[TSC_T_MIN_GAP];
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, want);
            const int slot = base + __popc(m & ((1u << tid) - 1u));
            if (want && slot >= Y.Vcap) { atomicOr(&c.h->err, ERR_OVERFLOW); want = false; }
            if (want) {
                const int hd = c.wq[s] + 1;
                const int at = c.sp_base[s] + hd;
                int4 nxt = make_int4(-1, INT_MAX, 0, 0);
                if (at < c.sp_base[s + S.n_spawn_lanes]) nxt = __ldg(S.spawn_rec + at);      // issued now, needed last
                c.pos[slot] = 0.0; c.spd[slot] = 0.0;
                c.rpos[slot] = rec.z;
                c.vid[slot] = rec.x; c.ellt[slot] = INT_MAX; c.blk[slot] = -1;
                c.dn[slot] = (u32) l | ((u32) (rec.w & 0xFFFF) << 16);
                c.pj[slot] = 0;
                c.foll[slot] = (u16) NONE16;
                if (n > 0) { const int t = c.tail[l]; c.lead[slot] = (u16) t; c.foll[t] = (u16) slot; }
                else { c.lead[slot] = (u16) NONE16; c.head[l] = (u16) slot; }
                c.tail[l] = (u16) slot;
                c.cnt[l] = (u16) (n + 1);
                c.wq[s] = (u16) hd;
                ++spawned;
                ((int4 *) c.sp_rec)[s] = nxt;
            }
            if (s < S.n_spawn_lanes) c.fresh[l] = want ? 1 : 0;
            base += __popc(m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) spawned += __shfl_xor_sync(0xffffffffu, spawned, o);
        if (tid == 0) {
            c.h->n_new = min(base, Y.Vcap); c.h->n_running += spawned;
            if (tick & 1) { c.h->n_ent = 0; c.h->n_x = 0; } else { c.h->n_h = 0; c.h->n_a = 0; }      // the NEXT tick's list counters
        }
        n_new_w0 = min(base, Y.Vcap);
        __syncwarp();
    }
    // ---- head vehicles (no vehicle ahead on their drivable) of the vehicles that were here before this tick, gathered
    //      warp-locally: look-ahead leader + gap (A.7).  Each WARP owns the slots of its 32-slot chunks.  When no spawn
    //      lane is fed by a lane-link (every shipped roadnet: vehicles start on roads that leave the network's rim), what
    //      the look-ahead reads -- the drivables AHEAD of a vehicle -- is never a lane handleWaiting touches, so it runs
    //      beside it; otherwise after a barrier. ----
    if (!S.spawn_pure) __syncthreads();
    {
        const int lane = tid & 31, w = tid >> 5, NW = NT / 32;
        const unsigned lt = (1u << lane) - 1u;
        const int n_chunks = (n_old + 31) >> 5;
        u16 *wl = c.xlist + w * Y.wl_cap;
        int nh = 0;
        for (int ch = w; ch < n_chunks; ch += NW) {
            const int i = (ch << 5) + lane;
            const bool hd = i < n_old && c.vid[i] >= 0 && c.lead[i] == NONE16;
            const unsigned m = __ballot_sync(0xffffffffu, hd);
            if (hd) wl[nh + __popc(m & lt)] = (u16) i;
            nh += __popc(m);
        }
        __syncwarp();
        for (int e = lane; e < nh; e += 32) {
            const int i = wl[e];
            int leader;
            double gap;
            head_look_ahead<ONE_T>(S, c, i, &leader, &gap);
            c.nblk[i] = (short) leader; c.npos[i] = gap;
        }
        __syncwarp();      // the chunks' owners read these in the decisions
    }
    if (!fused) __syncthreads();
    pt_mark(c, PT_SPAWN);
    int *const xctr = cross_counter(c, tick);

    // ---- getAction.  Every decision reads only the state the tick started from (positions, lists, blockers) and
    //      writes the vehicle's own entries of the next-state buffers, so the whole of it runs without a block-wide
    //      barrier: every vehicle, a lane each: car following (A.4); red light / blocked exit / turn speed (A.5 i-ii);
    //      then commit (finish_vehicle); the few vehicles that must examine the crosses of a lane-link (A.5 iii) are
    //      listed for the cross phase, which spreads each of them over a group of lanes. ----
    {
        const int lane = tid & 31, w = tid >> 5, NW = NT / 32;
        const unsigned lt = (1u << lane) - 1u;
        // fused: every warp its chunks of the vehicles that were here before the tick, then the first warp the new ones;
        // otherwise (after the barrier) the chunks cover the new ones as well
        const int n_lim = fused ? n_old : c.h->n_new;
        const int n_chunks = (n_lim + 31) >> 5;
        const int own = w < n_chunks ? (n_chunks - w + NW - 1) / NW : 0;
        const int extra = (fused && w == 0) ? (n_new_w0 - n_old + 31) >> 5 : 0;
        // (a vehicle that left a waiting buffer onto an empty lane this tick is a head too: its look-ahead runs here)
        for (int k = 0; k < own + extra; ++k) {
            const int i = k < own ? ((w + k * NW) << 5) + lane : n_old + ((k - own) << 5) + lane;
            const bool valid = k < own ? (i < n_lim && c.vid[i] >= 0) : (i < n_new_w0);
            const double *T = c.tmpl;
            u32 dnv = 0;
            int d = 0;
            double x = 0.0, v = 0.0, dlen = 0.0, ns = 0.0, vi = 0.0;
            bool zone = false, needx = false;
            if (valid) {
                T = tmpl_of<ONE_T>(S, c, c.vid[i]);
                dnv = c.dn[i];
                d = dnv & 0xFFFF;
                const int nd1 = (dnv >> 16) == 0xFFFFu ? -1 : (int) (dnv >> 16);
                x = c.pos[i]; v = c.spd[i];
                const double2 lm = __ldg(S.drv_lm + d);       // length, speed limit: one 16-byte load
                dlen = lm.x;
                int leader = c.lead[i];
                double gap;
                if (leader != (int) NONE16) gap = c.pos[leader] - tmpl_of<ONE_T>(S, c, c.vid[leader])[TSC_T_LEN] - x;
                else if (i >= n_old) head_look_ahead<ONE_T>(S, c, i, &leader, &gap);
                else { leader = c.nblk[i]; gap = c.npos[i]; }
                ns = T[TSC_T_MAX_SPEED];
                ns = min2(ns, v + T[TSC_T_MAX_POS_ACC] * dt);
                ns = min2(ns, lm.y);
                double cf = T[TSC_T_MAX_SPEED];
                if (leader >= 0) cf = car_follow_speed(T, v, gap, c.spd[leader], tmpl_of<ONE_T>(S, c, c.vid[leader])[TSC_T_MAX_NEG_ACC]);
                ns = min2(ns, cf);
                zone = d >= L || (nd1 >= L && dlen - x <= T[TSC_T_APPROACH_DIST]);   // intersection related speed applies (A.5)
                if (zone) {
                    vi = T[TSC_T_MAX_SPEED];
                    needx = true;
                    if (d < L) {
                        const int ll = nd1 - L;
                        const int el = __ldg(&S.llinfo[ll].end_lane);
                        bool enter = true;
                        if (c.cnt[el] > 0) {
                            int t = c.tail[el];
                            enter = c.pos[t] > tmpl_of<ONE_T>(S, c, c.vid[t])[TSC_T_LEN] + T[TSC_T_LEN] || c.spd[t] >= 2;
                        }
                        if (!ll_available(c, ll) || !enter) {
                            if (DIV_POS_HOT(0.5 * v * v, T[TSC_T_MAX_NEG_ACC]) > dlen - x) {
                                // cannot stop before the line any more
                            } else {
                                vi = min2(vi, stop_before_speed(T, v, dlen - x));
                                needx = false;      // stops at the line: the crosses are not examined
                            }
                        }
                        if (needx && __ldg(&S.llinfo[ll].type) != 3) vi = min2(vi, T[TSC_T_TURN_SPEED]);
                    }
                }
            }
            // vehicles that must examine the crosses of a lane-link (A.5 iii) go on the block's list: one ballot and
            // one atomic per warp; their car-following speed and the speed limit found so far wait in the next-state buffers
            const unsigned mx = __ballot_sync(0xffffffffu, needx);
            if (mx) {
                int base = 0;
                if (lane == 0) base = atomicAdd(xctr, __popc(mx));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (needx) { c.clist[base + __popc(mx & lt)] = (u16) i; c.nspd[i] = ns; c.npos[i] = vi; }
            }
            if (valid && !needx) {
                if (zone) ns = min2(ns, vi);
                finish_vehicle(S, Y, c, i, T, d, c.rpos[i], x, v, dlen, ns, -1);
            }
        }
    }
    __syncthreads();
    pt_mark(c, PT_PHASE1);
    const int n_slots = c.h->n_new;

    // ---- getAction, cross phase: Cross::canPass for every cross ahead of every listed vehicle.  canPass has no side
    //      effects, so all crosses are evaluated at once and the first refusal in link order is the sequential scan of
    //      A.5(iii): a group of G lanes (half / quarter / eighth of a warp: links rarely have more than a dozen crosses)
    //      per vehicle, one lane per cross, all listed vehicles at once; G = the widest group that still gives every
    //      listed vehicle its own group in one round ----
    {
        const int n_x = *xctr;
        const int G = n_x * 16 <= NT ? 16 : (n_x * 8 <= NT ? 8 : (n_x * 4 <= NT ? 4 : 2));
        const int lane = tid & 31, sl = lane & (G - 1);
        const unsigned gm = ((1u << G) - 1u) << (lane & ~(G - 1));
        for (int e = tid / G; e < n_x; e += NT / G) {
            const int i = c.clist[e];
            const double *T = tmpl_of<ONE_T>(S, c, c.vid[i]);
            const u32 dnv = c.dn[i];
            const int d = dnv & 0xFFFF;
            const bool on_ll = d >= L;
            const int ll = on_ll ? d - L : (int) (dnv >> 16) - L;
            const double x = c.pos[i], v = c.spd[i];
            const double dlen = __ldg(S.drv_length + d);
            const double dts = on_ll ? x : -(dlen - x);
            const int4 head = __ldg((const int4 *) &S.llinfo[ll]);   // start_lane, end_lane, cross_off, cross_end
            const int t1 = __ldg(&S.llinfo[ll].type);
            double vi = c.npos[i], ns = c.nspd[i];
            int blocker = -1;
            for (int base = head.z; base < head.w; base += G) {
                const int xi = base + sl;
                bool refuse = false;
                int foe = -1;
                double dOn = 0.0;
                if (xi < head.w) {
                    // (a per-tick "could this link announce a vehicle at all" bit per lane-link, tested before the entry is
                    // loaded, was measured: the cross phase got 1.2 k cycles per tick shorter, computing the bits cost 2.5 k)
                    CrossEntry X;
                    const int4 *src = (const int4 *) &S.cross[xi];
                    int4 *dst = (int4 *) &X;
                    dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);      // (measured slower: L1::no_allocate loads, 0.762 vs 0.742 ms;
                    // requesting the next round's entry before this one is evaluated, 0.706 vs 0.685)
                    dOn = X.dist;
                    if (!(dOn < dts)) refuse = !can_pass<ONE_T>(S, c, i, T, t1, X, dts, &foe);
                }
                const unsigned m = __ballot_sync(gm, refuse) & gm;
                if (m) {
                    const int src_lane = __ffs(m) - 1;
                    dOn = __shfl_sync(gm, dOn, src_lane);
                    foe = __shfl_sync(gm, foe, src_lane);
                    vi = min2(vi, stop_before_speed(T, v, dOn - dts - T[TSC_T_YIELD_DIST]));
                    blocker = foe;
                    break;
                }
            }
            ns = min2(ns, vi);
            if (sl == 0) finish_vehicle(S, Y, c, i, T, d, c.rpos[i], x, v, dlen, ns, blocker);
        }
    }
    // (the barrier also tells every thread, uniformly, whether any decision raised a sticky error)
    const bool bail = TICK_SYNC_ERR(c.h->err != 0);
#ifdef TSC_PHASE_TIMING
    if (c.pt && tid == 0) atomicAdd(c.pt + PT_NX, (unsigned long long) *xctr);
#endif
    pt_mark(c, PT_PHASE2);

    // ---- updateLocation.  Slots are stable, so there is nothing to re-pack: every vehicle's next state is
    //      already in the next-state buffers, and only the movers of the tick (a few dozen) touch the lists.
    //      Leaving: movers are a prefix of their drivable's list (FIFO); the mover at its head walks the prefix
    //      and hands the head over.  Entering (after a barrier): entrants go behind the vehicles that stay,
    //      ordered by new distance (descending, ties by creation id); one thread per entered drivable. ----
    const int n_mv = *mover_counter(c, tick);      // (stable: the next tick counts in the other pair)
    if (bail) {   // mover list overflow / no slot left: keep the old state; the sticky flag reports it
        __syncthreads();
        if (tid == 0) { c.h->tick = tick + 1; c.h->n_slots = n_slots; }      // (vehicles that did enter before the slots ran out stay listed)
        __syncthreads();
        return true;
    }
    for (int m = tid; m < n_mv; m += NT) {
        const int i = c.mv_slot[m];
        const int d = c.dn[i] & 0xFFFF;
        const int ld = c.lead[i];
        if (ld == (int) NONE16) {      // head of its drivable: hand the head over to the first vehicle that stays
            int k = 0, v = i;
            while (v != (int) NONE16 && (c.pj[v] & PJ_MOVER) && k < Y.Vcap) { ++k; v = c.foll[v]; }
            c.head[d] = (u16) v;
            if (v != (int) NONE16) c.lead[v] = (u16) NONE16; else c.tail[d] = (u16) NONE16;
            if (k != (int) c.leave[d]) atomicOr(&c.h->err, ERR_ORDER);
            c.cnt[d] = (u16) (c.cnt[d] - k);
            c.leave[d] = 0;
        } else if (!(c.pj[ld] & PJ_MOVER)) atomicOr(&c.h->err, ERR_ORDER);      // left although the vehicle ahead stays
        if (c.mv_to[m] == NONE16) {    // finished: statistics (A.8); the slot becomes a hole
            const int vt = __ldg(S.veh_tick + c.vid[i]);
            atomicAdd((unsigned long long *) &c.h->cum_tt, (unsigned long long) (tick - vt));
            atomicAdd((unsigned long long *) &c.h->fin_enter, (unsigned long long) vt);
            atomicAdd(&c.h->n_finished, 1);
            atomicSub(&c.h->n_running, 1);
            c.vid[i] = -1; c.nblk[i] = -1;
        }
    }
    // up to 32 movers (the usual case: half a dozen) are all lanes of the first warp: a warp-level barrier orders the
    // two halves of the surgery and the other warps go straight to the end of the tick
    if (Y.warp_surgery && n_mv <= 32) { if (tid < 32) __syncwarp(); } else __syncthreads();
    pt_mark(c, PT_LEAVE);
    for (int m = tid; m < n_mv; m += NT) {
        const int i = c.mv_slot[m];
        const int dd = c.mv_to[m];
        if (dd == (int) NONE16) { c.pj[i] = 0; continue; }
        // the vehicle's own fields
        const int q = c.mv_q[m];
        c.rpos[i] = q;
        c.dn[i] = (u32) dd | ((u32) (__ldg(S.route_seq + q + 1) & 0xFFFF) << 16);
        c.pj[i] = c.mv_pj[m];
        c.ellt[i] = dd >= L ? tick : INT_MAX;      // enterLaneLinkTime: the tick a lane-link was entered, "never" on a lane
        // the list of the drivable it enters
        const int n_in = c.ent[dd];
        if (n_in == 1) {
            const int t = c.cnt[dd] > 0 ? (int) c.tail[dd] : (int) NONE16;
            c.lead[i] = (u16) t; c.foll[i] = (u16) NONE16;
            if (t != (int) NONE16) c.foll[t] = (u16) i; else c.head[dd] = (u16) i;
            c.tail[dd] = (u16) i;
            c.cnt[dd] = (u16) (c.cnt[dd] + 1);
            c.ent[dd] = 0;
        } else {
            // several entrants: the one listed first appends them all, in order of new distance
            bool first = true;
            for (int k = 0; k < m; ++k) if (c.mv_to[k] == dd) { first = false; break; }
            if (!first) continue;
            int t = c.cnt[dd] > 0 ? (int) c.tail[dd] : (int) NONE16;
            double last_x = 0.0;
            int last_v = 0;
            for (int r = 0; r < n_in; ++r) {      // selection in (distance desc, creation id asc) order
                int best = -1;
                double bx = 0.0;
                int bv = 0;
                for (int k = m; k < n_mv; ++k) {
                    if (c.mv_to[k] != dd) continue;
                    const int o = c.mv_slot[k];
                    const double ox = c.npos[o];
                    const int ov = c.vid[o];
                    if (r > 0 && !(ox < last_x || (ox == last_x && ov > last_v))) continue;      // already placed
                    if (best < 0 || ox > bx || (ox == bx && ov < bv)) { best = o; bx = ox; bv = ov; }
                }
                if (best < 0) break;
                c.lead[best] = (u16) t; c.foll[best] = (u16) NONE16;
                if (t != (int) NONE16) c.foll[t] = (u16) best; else c.head[dd] = (u16) best;
                t = best; last_x = bx; last_v = bv;
            }
            c.tail[dd] = (u16) t;
            c.cnt[dd] = (u16) (c.cnt[dd] + n_in);
            c.ent[dd] = 0;
        }
    }
    if (tid == 0) {
#ifdef TSC_PHASE_TIMING
        if (c.pt) atomicAdd(c.pt + PT_NENT, (unsigned long long) n_mv);
#endif
        c.h->n_slots = n_slots; c.h->tick = tick + 1;
    }
    // the next state becomes the current one
    c.par ^= 1;
    set_pingpong(Y, c);
    const bool err_now = TICK_SYNC_ERR(c.h->err != 0);      // (the list surgery may have found a FIFO violation)
    pt_mark(c, PT_ENTER);
    return err_now;
}

// ----------------------------------------------------------------------------
// pytsc layer: phase program, Retriever, per-signal stats, reward, mask, obs
// ----------------------------------------------------------------------------
// float("%f" % x): the reference reads distance and speed through
// get_vehicle_info strings (retriever.py:35-36,44) -- six decimals, correctly
// rounded both ways.  x >= 0.
__device__ double round6(double x) {
    double p = x * 1e6;
    double e = __fma_rn(x, 1e6, -p);       // exact product = p + e
    double n = rint(p);
    double dlt = p - n;
    if (dlt == 0.5 && e > 0) n += 1.0;
    else if (dlt == -0.5 && e < 0) n -= 1.0;
    return n / 1e6;
}

// Python's float floor division a // b (floatobject.c float_divmod), b > 0.
__device__ double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0) != (mod < 0))) div -= 1.0;
    if (div != 0.0) {
        double f = floor(div);
        if (div - f > 0.5) f += 1.0;
        return f;
    }
    return 0.0;
}

// numpy's pairwise float64 sum for n <= 128 (8 interleaved accumulators).
__device__ double np_sum(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// ---- Retriever._compute_lane_position_matrix (retriever.py:20-52) seen through a window: the last
//      (incoming side, traffic_signal.py:124) or first (outgoing side, :135) `vis` bins of lane l's
//      padded list.  w[k] = -1 for an empty bin, else -1 + sum over its vehicles of (1 + speed / max). ----
__device__ void lane_window(const DevScn &S, const Ctx &c, int l, bool tail, double *w) {
    const int vis = S.visibility;
    const double plen = __ldg(S.lane_pytsc_length + l);
    const double mspeed = __ldg(S.drv_max_speed + l);
    const int bins = (int) (plen / S.v_size);
    const int n = c.cnt[l];
    for (int k = 0; k < vis; ++k) w[k] = -1.0;
    if (bins > 0 && n > 0) {
        const int len = bins < vis ? vis : bins;     // padded length
        const int lo = tail ? len - vis : 0;         // window start in the padded list
        const double bin_size = plen / bins;
        int v = c.head[l];
        for (int k = 0; k < n && v != (int) NONE16; ++k, v = c.foll[v]) {      // the lane's vehicles, front to back
            double p = round6(c.pos[v]);
            if (p < 0) p = 0; else if (p > plen) p = plen;
            int bi = trunc_int_x86(py_floordiv(p, bin_size));
            if (bi >= bins) bi = bins - 1;
            const int wi = bi - lo;
            if (wi >= 0 && wi < vis) {
                const double nsp = round6(c.spd[v]) / mspeed;
                w[wi] += 1.0;
                w[wi] += nsp;
            }
        }
    }
}

// ---- get_allowable_phase_switches (common/traffic_signal.py:329-361 free, 375-404 round robin):
//      bit p = pytsc phase index p may be selected now ----
__device__ u32 allowable_phases(const DevScn &S, const Ctx &c, int s) {
    const int cur = c.scur[s], t = c.stop[s], P = __ldg(S.sig_n_phases + s);
    u32 allow = 0;
    if (__ldg(S.sig_phase_green + s * S.P + cur)) {
        const int mn = __ldg(S.sig_min_time + s * S.P + cur), mx = __ldg(S.sig_max_time + s * S.P + cur);
        const int nxt = (cur + 1) % P;
        if (t < mn) allow = 1u << cur;
        else if (t < mx) allow = (1u << cur) | (1u << nxt);
        else if (t == mx) allow = 1u << nxt;
    } else if (S.round_robin) {
        allow = 1u << ((cur + 1) % P);
    } else {
        for (int p = 0; p < P; ++p)
            if (__ldg(S.sig_phase_green + s * S.P + p) && p != cur - 1) allow |= 1u << p;
    }
    return allow;
}

// counter-based generator for tie breaks (the reference draws np.random.choice): splitmix64 of
// (seed, replica, signal, tick)
__device__ __forceinline__ u32 ctl_random(int seed, int b, int s, int tick) {
    unsigned long long z = (unsigned long long) (u32) seed * 0x9E3779B97F4A7C15ull + ((unsigned long long) (u32) b << 32 | (u32) s);
    z ^= (unsigned long long) (u32) tick * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (u32) (z >> 16);
}

// ---- pytsc's rule-based controllers (controllers/controllers.py:57-268) for every signal of the
//      replica, from the state as it is now: decided[s] = the pytsc phase index get_action returns ----
template <int NT>
__device__ void controller_decide(const DevScn &S, Ctx &c, const StepArgs &a, int b, int *decided) {
    const int mode = a.apply_actions, L = S.L, A = S.A, vis = S.visibility;
    // per lane: bins of the incoming-side window holding exactly one standing vehicle (== 0.0,
    // controllers.py:111), holding any vehicle (>= 0.0, :163, :236), and the same on the outgoing side (:168)
    u8 *tail_zero = (u8 *) c.npos, *tail_any = tail_zero + L, *head_any = tail_any + L;
    if (mode != 7) {
        for (int l = threadIdx.x; l < L; l += NT) {
            int z = 0, n = 0, h = 0;
            if (c.cnt[l] > 0) {
                double w[16];
                lane_window(S, c, l, true, w);
                for (int k = 0; k < vis; ++k) { z += w[k] == 0.0; n += w[k] >= 0.0; }
                lane_window(S, c, l, false, w);
                for (int k = 0; k < vis; ++k) h += w[k] >= 0.0;
            }
            tail_zero[l] = (u8) z; tail_any[l] = (u8) n; head_any[l] = (u8) h;
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < A; s += NT) {
        const int P = __ldg(S.sig_n_phases + s), cur = c.scur[s], t = c.stop[s];
        const int nxt = (cur + 1) % P;
        const bool green = __ldg(S.sig_phase_green + s * S.P + cur) != 0;
        const u32 allow = allowable_phases(S, c, s);
        int *sc = a.ctl_scores ? a.ctl_scores + ((size_t) b * A + s) * S.P : nullptr;
        if (sc) for (int p = 0; p < S.P; ++p) sc[p] = (mode == 4 || mode == 5) ? INT_MIN : 0;
        int idx = nxt;
        if (mode == 2) {            // FixedTimeController (controllers.py:39-54)
            idx = (green && t < a.controller_arg) ? cur : nxt;
        } else if (mode == 4 || mode == 5) {   // Greedy (:66-114) / MaxPressure (:123-176)
            if (green) {
                int best = INT_MIN, n_tie = 0;
                for (int pass = 0; pass < 2; ++pass) {
                    const int pick = pass ? (int) (ctl_random(a.controller_arg, b, s, c.h->tick) % (u32) (n_tie > 0 ? n_tie : 1)) : 0;
                    int seen = 0;
                    for (int p = 0; p < P; ++p) {
                        if (!((allow >> p) & 1u)) continue;
                        int score = 0;
                        for (int e = __ldg(S.ctl_off + s * S.P + p), e1 = __ldg(S.ctl_off + s * S.P + p + 1); e < e1; ++e) {
                            const int li = __ldg(S.ctl_in_lane + e);
                            if (mode == 4) score += tail_zero[li];
                            else {
                                const int lo = __ldg(S.ctl_out_lane + e);
                                const int d = (int) tail_any[li] - (lo >= 0 ? (int) head_any[lo] : 0);
                                score += d < 0 ? -d : d;
                            }
                        }
                        if (!pass) {
                            if (sc) sc[p] = score;
                            if (score > best) { best = score; n_tie = 1; } else if (score == best) ++n_tie;
                        } else if (score == best) {
                            if (seen == pick) { idx = p; break; }
                            ++seen;
                        }
                    }
                }
            }
        } else if (mode == 6) {     // SOTL (:199-238)
            if ((allow >> cur) & 1u) {
                const int theta = a.controller_arg & 0xFF, mu = (a.controller_arg >> 8) & 0xFF, phi_min = (a.controller_arg >> 16) & 0xFFFF;
                int flow[2];
                for (int q = 0; q < 2; ++q) {
                    const int p = q ? (cur + 2) % P : cur;      // next_green_phase_index (common/traffic_signal.py:228-238)
                    int f = 0;
                    for (int e = __ldg(S.ctl_off + s * S.P + p), e1 = __ldg(S.ctl_off + s * S.P + p + 1); e < e1; ++e)
                        f += tail_any[__ldg(S.ctl_in_lane + e)];
                    flow[q] = f;
                }
                if (sc) { sc[0] = flow[0]; if (S.P > 1) sc[1] = flow[1]; }
                idx = (t >= phi_min && !(0 < flow[0] && flow[0] < mu) && flow[1] >= theta) ? nxt : cur;
            }
        } else {                    // Random (:252-268): uniform over the allowed phase indices
            const int n = __popc(allow);
            int pick = n ? (int) (ctl_random(a.controller_arg, b, s, c.h->tick) % (u32) n) : 0;
            for (int p = 0; p < P; ++p)
                if ((allow >> p) & 1u) { if (pick == 0) { idx = p; break; } --pick; }
        }
        decided[s] = idx;
        if (a.ctl_actions) a.ctl_actions[(size_t) b * A + s] = idx;
    }
    __syncthreads();
}

template <int NT>
__device__ void apply_controller(const DevScn &S, Ctx &c, const StepArgs &a, int b, const int *decided) {
    for (int s = threadIdx.x; s < S.A; s += NT) {
        if (a.set_raw_phase) {
            const int r = a.raw_phase[(size_t) b * S.A + s];
            if (r < 0 || r >= __ldg(S.sig_n_raw + s)) atomicOr(&c.h->err, ERR_BAD_PHASE);
            else c.sraw[s] = (u8) r;
        }
        if (a.init_program >= 0) {
            c.scur[s] = (u8) a.init_program; c.schg[s] = 0; c.stop[s] = 0;
            c.sraw[s] = (u8) __ldg(S.sig_phase_raw + s * S.P + a.init_program);
        }
        if (a.apply_actions) {
            int P = __ldg(S.sig_n_phases + s);
            int cur = c.scur[s], t = c.stop[s];
            int idx;
            if (a.apply_actions == 2) {   // FixedTimeController.get_action (controllers.py:39-54)
                idx = (__ldg(S.sig_phase_green + s * S.P + cur) && t < a.controller_arg) ? cur : (cur + 1) % P;
            } else if (a.apply_actions >= 4) {   // the rule-based controller's choice (controller_decide)
                idx = decided[s];
                if (idx < 0 || idx >= P) idx = cur;
            } else {
                int act = a.actions[(size_t) b * S.A + s];
                if (a.apply_actions == 3) idx = act;                                               // TSController.switch_phase(index)
                else if (S.action_space == TSC_ACT_PHASE_SWITCH) idx = act == 1 ? (cur + 1) % P : cur;   // actions.py:152-158
                else idx = act;                                                                    // actions.py:106-108
                if (idx < 0 || idx >= P) idx = cur;
            }
            // BaseTSProgram.update_current_phase (common/traffic_signal.py:94-109)
            if (idx == cur) { c.schg[s] = 0; t += S.yellow_time; }
            else { c.schg[s] = 1; t = S.yellow_time; }
            c.scur[s] = (u8) idx; c.stop[s] = t;
            c.sraw[s] = (u8) __ldg(S.sig_phase_raw + s * S.P + idx);   // engine.set_tl_phase (traffic_signal.py:58)
        }
    }
}

__device__ __forceinline__ float ref_trunc(double x, bool exact) { return (float) (exact ? trunc(x) : x); }

template <int NT>
__device__ void retrieve(const DevScn &S, const Layout &Y, Ctx &c, const StepArgs &a, int b, unsigned char *smem) {
    const int tid = threadIdx.x;
    const int L = S.L, A = S.A;
    const tsc_outputs_t &O = a.out;
    // scratch: the next-state kinematics pair (npos | nspd, contiguous, 16 bytes per laid-out slot) is free between ticks
    double *l_occ = c.npos;              // [L]
    double *l_ms = l_occ + L;            // [L]
    double *l_nms = l_ms + L;            // [L] mean speed / lane speed limit (metrics.py:113-135, traffic_signal.py:118)
    double *s_loc = l_nms + L;           // [A] local reward term
    double *s_prs = s_loc + A;           // [A] pressure
    int *l_q = (int *) (s_prs + A);      // [L]
    // what follows, 16-byte aligned: position-matrix windows, or the host packet
    double *const r_tail = (double *) ((unsigned char *) c.npos + ((24 * L + 16 * A + 4 * L + 15) & ~15));
    // registered host path: the replica's packet is assembled here, then stored to host memory in 16-byte pieces
    unsigned char *const pkst = a.pk ? (unsigned char *) r_tail : nullptr;

    // --- Retriever._compute_lane_measurements (retriever.py:54-85) ---
    for (int l = tid; l < L; l += NT) {
        const int n = c.cnt[l];
        int q = 0;
        double tot = 0.0;
        int v = n > 0 ? (int) c.head[l] : (int) NONE16;
        for (int k = 0; k < n && v != (int) NONE16; ++k) {      // front to back: the sum keeps the reference's order
            const double sp = c.spd[v];
            v = c.foll[v];
            tot += sp;
            q += sp < 0.1;
        }
        double ms = n ? div_pos(tot, (double) n) : 0.0;
        double occ = div_pos((double) n, __ldg(S.lane_cells + l));      // lane_cells = pytsc lane length / veh_size_min_gap
        l_occ[l] = occ; l_ms[l] = ms; l_q[l] = q;
        l_nms[l] = div_pos(ms, __ldg(S.drv_max_speed + l));
        size_t o = (size_t) b * L + l;
        if (O.lane_count) O.lane_count[o] = n;
        if (O.lane_queued) O.lane_queued[o] = q;
        if (O.lane_occupancy) O.lane_occupancy[o] = (float) occ;
        if (O.lane_mean_speed) O.lane_mean_speed[o] = (float) ms;
        if (O.lane_meas64) { O.lane_meas64[2 * o] = occ; O.lane_meas64[2 * o + 1] = ms; }
    }
    __syncthreads();
    pt_mark(c, PT_PHASE1A);      // (debug timing: retrieve, lane sums)
    if (pkst) {      // per incoming lane, in observation-row order, the three values the row shows (observations.py:313-321)
        for (int i = tid; i < S.n_in_total; i += NT) {
            const u32 e = __ldg(S.pk_lane + i);
            const int l = (int) (e & 0x7FFFFFFFu);
            const bool tr = (e >> 31) != 0;
            if (S.pk_mode) {
                const u32 q = (u32) min(l_q[l], 255), oc = (u32) min((int) trunc(l_occ[l]), 255), ms = (u32) min((int) trunc(l_ms[l]), 255);
                ((u32 *) pkst)[i] = q | (oc << 8) | (ms << 16);
            } else {
                float *f = (float *) pkst + 3 * i;
                f[0] = (float) l_q[l]; f[1] = ref_trunc(l_occ[l], tr); f[2] = ref_trunc(l_ms[l], tr);
            }
        }
    }

    // --- MetricsParser.density_map (backends/cityflow/metrics.py:170-199): thread per signal pair ---
    if (O.density_map) {
        auto directed = [&](int i, int j) -> double {
            const int e0 = __ldg(S.dm_off + i * A + j), e1 = __ldg(S.dm_off + i * A + j + 1);
            if (e1 == e0) return 0.0;
            double tot = 0.0;
            for (int e = e0; e < e1; ++e) tot += l_occ[__ldg(S.dm_lane + e)];
            const double m = tot / (double) (e1 - e0);
            return m < 0.0 ? 0.0 : (m > 1.0 ? 1.0 : m);
        };
        double *dm = O.density_map + (size_t) b * A * A;
        for (int p = tid; p < A * A; p += NT) {
            const int i = p / A, j = p - i * A;
            dm[p] = (directed(i, j) + directed(j, i)) / 2 + 1e-6 * __ldg(S.dm_adjacency + p);
        }
    }

    // --- position-matrix windows (retriever.py:20-52, traffic_signal.py:124,135) ---
    const int vis = S.visibility;
    const bool need_pos = O.pos_in || O.pos_out || (O.obs && S.obs_type == TSC_OBS_POSITION_MATRIX);
    double *win_in = r_tail;               // [n_in_total][vis] scratch (fp64), only when needed
    if (need_pos) {
        const int n_in = S.n_in_total, n_out = S.n_out_total;
        for (int e = tid; e < n_in + n_out; e += NT) {
            bool inc = e < n_in;
            int l = inc ? __ldg(S.sig_in_lane + e) : __ldg(S.sig_out_lane + e - n_in);
            double w[16];
            lane_window(S, c, l, inc, w);
            if (inc) {
                if (S.obs_type == TSC_OBS_POSITION_MATRIX) for (int k = 0; k < vis; ++k) win_in[e * vis + k] = w[k];
                if (O.pos_in) for (int k = 0; k < vis; ++k) O.pos_in[((size_t) b * n_in + e) * vis + k] = (float) w[k];
            } else if (O.pos_out) {
                for (int k = 0; k < vis; ++k) O.pos_out[((size_t) b * n_out + (e - n_in)) * vis + k] = (float) w[k];
            }
        }
        __syncthreads();
    }

    // --- TrafficSignal.update_stats (traffic_signal.py:101-141), local reward term, mask ---
    for (int s = tid; s < A; s += NT) {
        int i0 = __ldg(S.sig_in_off + s), i1 = __ldg(S.sig_in_off + s + 1);
        int o0 = __ldg(S.sig_out_off + s), o1 = __ldg(S.sig_out_off + s + 1);
        int nq = 0;
        double occ = 0, ms = 0, md = 0, oocc = 0;
        for (int e = i0; e < i1; ++e) {
            int l = __ldg(S.sig_in_lane + e);
            nq += l_q[l];
            occ += l_occ[l];
            ms += l_ms[l];
            md += 1 - l_nms[l];
        }
        int nin = i1 - i0, nout = o1 - o0;
        occ /= nin; ms /= nin; md /= nin;
        for (int e = o0; e < o1; ++e) oocc += l_occ[__ldg(S.sig_out_lane + e)];
        oocc /= nout;
        double pressure = fabs(occ - oocc);
        int cur = c.scur[s], t = c.stop[s], P = __ldg(S.sig_n_phases + s);
        double ntop = (double) t / (double) __ldg(S.sig_max_time + s * S.P + cur);
        if (O.sig_stats64) {
            double *o = O.sig_stats64 + ((size_t) b * A + s) * 8;
            o[0] = nq; o[1] = occ; o[2] = ms; o[3] = md; o[4] = oocc; o[5] = pressure; o[6] = ntop; o[7] = cur;
        }
        // local reward term (reward.py:77-80 | 125-128)
        double chg = c.schg[s] ? 1.0 : 0.0;
        double metric = S.reward_type == TSC_REWARD_QUEUE ? (double) nq : pressure;
        s_loc[s] = -S.flick * chg - metric - 1e-6;
        s_prs[s] = pressure;

        // action mask (common/traffic_signal.py:329-361, 375-404; actions.py:119-131, 169-188)
        if (O.mask || pkst) {
            const u32 allow = allowable_phases(S, c, s);
            u32 bits;      // bit k = action k allowed
            if (S.action_space == TSC_ACT_PHASE_SWITCH) bits = ((allow >> cur) & 1u) | (((allow >> ((cur + 1) % P)) & 1u) << 1);
            else bits = P >= 32 ? allow : (allow & ((1u << P) - 1u));
            if (O.mask) {
                u8 *m = O.mask + ((size_t) b * A + s) * S.n_actions;
                for (int p = 0; p < S.n_actions; ++p) m[p] = (u8) ((bits >> p) & 1u);
            }
            if (pkst) { pkst[S.pk_o_phase + s] = (u8) cur; ((u32 *) (pkst + S.pk_o_mask))[s] = bits; }
        }
    }
    __syncthreads();
    pt_mark(c, PT_PHASE1C);      // (debug timing: retrieve, per-signal block)

    // --- position-matrix observation rows (observations.py:140-160; :72-88 drops window entries <= 0, so a lane's block has
    //     variable length and later lanes shift left): a thread per incoming lane.  Each counts the entries of the lanes
    //     before it in its signal's row (at most 15 short sums), then writes its own block; a thread per signal pads the row
    //     and appends the phase one-hot. ---
    if (O.obs && S.obs_type == TSC_OBS_POSITION_MATRIX) {
        const int ML = S.max_lanes_per_signal, MP = S.max_obs_phases;
        const bool ex = S.reference_exact != 0;
        const int body = ML * (vis + 9);
        auto lane_entries = [&](int e) { int n = 9; for (int k = 0; k < vis; ++k) n += win_in[e * vis + k] > 0; return n; };
        for (int e = tid; e < S.n_in_total; e += NT) {
            const int sg = __ldg(S.in_sig + e);
            const int i0 = __ldg(S.sig_in_off + sg), i1 = __ldg(S.sig_in_off + sg + 1);
            int before = 0, total = 0;
            for (int q = i0; q < i1; ++q) { const int n = lane_entries(q); total += n; if (q < e) before += n; }
            const bool tr = ex && total < body;      // pad_list truncates only if it pads
            float *dst = O.obs + ((size_t) b * A + sg) * S.obs_dim;
            const int l = __ldg(S.sig_in_lane + e);
            int k = before;
            for (int f = 0; f < 9 && k < body; ++f) dst[k++] = ref_trunc(__ldg(S.lane_feat + l * 9 + f), tr);
            for (int j = 0; j < vis; ++j) {
                const double val = win_in[e * vis + j];
                if (val > 0 && k < body) dst[k++] = ref_trunc(val > 1.0 ? 1.0 : val, tr);   // np.clip(val + 0, 0, 1)
            }
        }
        for (int sg = tid; sg < A; sg += NT) {
            const int i0 = __ldg(S.sig_in_off + sg), i1 = __ldg(S.sig_in_off + sg + 1);
            int k = 0;
            for (int q = i0; q < i1; ++q) k += lane_entries(q);
            float *dst = O.obs + ((size_t) b * A + sg) * S.obs_dim;
            if (k > body) k = body;
            for (; k < body; ++k) dst[k] = -1.0f;
            const int cur = c.scur[sg], P = __ldg(S.sig_n_phases + sg);
            for (int p = 0; p < MP; ++p) dst[k++] = p < P ? (p == cur ? 1.0f : 0.0f) : -1.0f;
        }
    }

    // --- local rewards with spatially discounted neighbours (reward.py:81-88 | 129-136) ---
    if (O.reward || pkst) {
        for (int s = tid; s < A; s += NT) {
            double r = s_loc[s];
            int n0 = __ldg(S.nbr_off + s), n1 = __ldg(S.nbr_off + s + 1);
            for (int e = n0; e < n1; ++e) r += __ldg(S.nbr_weight + e) * s_loc[__ldg(S.nbr_idx + e)];
            if (O.reward) O.reward[(size_t) b * A + s] = (float) r;
            if (pkst) ((float *) (pkst + S.pk_o_reward))[s] = (float) r;
        }
    }

    // --- network metrics (metrics.py) and the global reward: the last warp (the first ones own the signals) ---
    if (tid >= NT - 32) {
        const int ln = tid & 31;
        int qsum = 0, vsum = 0;
        double wspeed = 0, occs = 0, nms = 0;
        for (int l = ln; l < L; l += 32) {
            int n = c.cnt[l];
            qsum += l_q[l]; vsum += n;
            wspeed += l_ms[l] * n;
            occs += l_occ[l];
            nms += l_nms[l];
        }
        int chg = 0;
        for (int s = ln; s < A; s += 32) chg += c.schg[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
            vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
            chg += __shfl_xor_sync(0xffffffffu, chg, o);
            wspeed += __shfl_xor_sync(0xffffffffu, wspeed, o);
            occs += __shfl_xor_sync(0xffffffffu, occs, o);
            nms += __shfl_xor_sync(0xffffffffu, nms, o);
        }
        if (ln == 0) {
            double flicker = (double) chg / (double) A;
            double psum = A <= 128 ? np_sum(s_prs, A) : 0.0;
            if (A > 128) for (int s = 0; s < A; ++s) psum += s_prs[s];
            double density = occs / L, norm_ms = nms / L;
            if (O.metrics) {
                double *m = O.metrics + (size_t) b * 8;
                m[0] = qsum; m[1] = vsum ? div_pos(wspeed, (double) vsum) : 0.0; m[2] = 1 - norm_ms; m[3] = density;
                m[4] = psum; m[5] = density * norm_ms; m[6] = flicker; m[7] = norm_ms;
            }
            if (O.reward_global || pkst) {
                double r;
                if (S.reward_type == TSC_REWARD_QUEUE) { r = 1e-6; r += S.flick * flicker; r += qsum; r = -1 * r; }
                else { r = 1e-6; r -= S.flick * flicker; r -= psum; }
                if (O.reward_global) O.reward_global[b] = (float) r;
                if (pkst) *(float *) (pkst + S.pk_o_rg) = (float) r;
            }
            if (O.err) O.err[b] = (int) c.h->err;
            if (O.sim) {
                int now = c.h->tick;
                int tt = now < S.horizon + 1 ? now : S.horizon + 1;
                const size_t fo = (size_t) c.h->flow_set * (S.horizon + 2) + tt;
                long long created = __ldg(S.created_cnt + fo);
                long long alive = created - c.h->n_finished;
                double total = (double) (c.h->cum_tt + alive * now - (__ldg(S.created_enter + fo) - c.h->fin_enter)) * S.interval;
                long long n = c.h->n_finished + alive;
                double *o = O.sim + (size_t) b * 4;
                o[0] = c.h->n_running; o[1] = n == 0 ? 0.0 : total / (double) n; o[2] = now * S.interval; o[3] = c.h->n_finished;
            }
        }
    }

    // --- lane-feature observation / state rows (observations.py:305-329, 352-374): one element per
    //     thread, rows are contiguous so the stores coalesce.  What each element is was worked out once
    //     on the host (build_obs_recipe): a static value (lane features, padding) or a code naming the
    //     dynamic quantity to read ---
    {
        const int n_el = A * S.state_dim;
        float *obs = (O.obs && S.obs_type == TSC_OBS_LANE_FEATURES) ? O.obs + (size_t) b * n_el : nullptr;
        float *state = O.state ? O.state + (size_t) b * n_el : nullptr;
        if (obs || state) {
            auto element = [&](u32 code, float val) -> float {
                if (code) {
                    const int kind = code & 7, arg = (int) (code >> 4);
                    const bool tr = (code & 8) != 0;          // pad_list only converts when it pads
                    if (kind == 1) val = (float) l_q[arg];
                    else if (kind == 2) val = ref_trunc(l_occ[arg], tr);
                    else if (kind == 3) val = ref_trunc(l_ms[arg], tr);
                    else val = (arg & 0xFF) == c.scur[arg >> 8] ? 1.0f : 0.0f;      // phase one-hot
                }
                return val;
            };
            const bool vec = (n_el & 3) == 0 && ((((size_t) obs) | ((size_t) state)) & 15) == 0;
            if (vec) {      // four consecutive elements per thread: 16-byte loads of the recipe, 16-byte row stores
                const uint4 *code4 = (const uint4 *) S.obs_code;
                const float4 *stat4 = (const float4 *) S.obs_static;
                for (int q = tid; q < n_el / 4; q += NT) {
                    const uint4 cd = __ldg(code4 + q);
                    float4 v = __ldg(stat4 + q);
                    v.x = element(cd.x, v.x); v.y = element(cd.y, v.y); v.z = element(cd.z, v.z); v.w = element(cd.w, v.w);
                    if (obs) ((float4 *) obs)[q] = v;
                    if (state) ((float4 *) state)[q] = v;
                }
            } else {
                for (int idx = tid; idx < n_el; idx += NT) {
                    const float val = element(__ldg(S.obs_code + idx), __ldg(S.obs_static + idx));
                    if (obs) obs[idx] = val;
                    if (state) state[idx] = val;
                }
            }
        }
    }
    // ---- registered host path: the packet goes straight to page-locked host memory (coalesced 16-byte
    //      stores over PCIe), then a flag is raised behind a system-scope fence; host threads follow the
    //      flags while the other replicas of the launch are still being stepped ----
    if (pkst) {
        __syncthreads();
        uint4 *dst = (uint4 *) (a.pk + (size_t) b * S.pk_bytes);
        const uint4 *src = (const uint4 *) pkst;
        for (int i = tid; i < S.pk_bytes / 16; i += NT) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            *(volatile u32 *) (a.pk_flags + b) = a.pk_seq;
        }
    }
    // no trailing barrier: nothing below writes what the slower warps still read (the caller
    // synchronises before the next replica is staged in)
}

// ----------------------------------------------------------------------------
// The step kernel
// ----------------------------------------------------------------------------
__device__ __forceinline__ void copy16(void *dst, const void *src, int bytes, int tid, int nt) {
    // bytes is a multiple of 16, both pointers 16-byte aligned
    const int4 *s = (const int4 *) src;
    int4 *d = (int4 *) dst;
    for (int i = tid; i < bytes / 16; i += nt) d[i] = s[i];
}

// The same copy as cp.async (global -> shared, 16 bytes per request, no register staging): every request of
// every column is in flight before the first one is waited for.
__device__ __forceinline__ void copy16_async(void *dst_smem, const void *src, int bytes, int tid, int nt) {
    const unsigned d0 = (unsigned) __cvta_generic_to_shared(dst_smem);
    const char *s = (const char *) src;
    for (int i = tid * 16; i < bytes; i += nt * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i), "l"(s + i) : "memory");
}
__device__ __forceinline__ void copy16_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// ---- bulk asynchronous copies (the TMA engine's 1-D form): ONE thread issues a whole image column per instruction; an
//      mbarrier in shared memory counts the bytes that have landed.  Sizes and addresses are multiples of 16 bytes. ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
    if (bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src),
                     "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, unsigned bytes) {
    if (bytes) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// CTL: the rule-based controllers are compiled in (kept out of the plain variant: their code costs the
// hot path 2-3 % through register allocation alone).
// GMEM: the replica's working set does not fit an SM's shared memory (a 16 x 16 grid needs ~1 MB): the
// block works out of a global-memory workspace instead -- same layout, same code, L2-resident.
// VC > 0: built for exactly VC vehicle slots -- the per-vehicle columns sit at compile-time offsets of the working set
// (v_offsets), so their addresses are immediates of the shared-memory instructions instead of a constant-bank load and
// an add at every use (a tenth of the executed instructions of the generic build, and registers it could not spare).
template <int NT, int MINB, bool CTL, bool GMEM, bool ONE_T, int VC = 0>
__global__ void __launch_bounds__(NT, MINB) tsc_step_kernel(const DevScn S, const Layout Yparam, unsigned char *images, const StepArgs a) {
    Layout Yl = Yparam;
    if (VC > 0) {
        constexpr VOff v = v_offsets(VC > 0 ? VC : 32, VC > 0 ? VC : 32, NT / 32);
        Yl.Vcap = VC; Yl.Vlay = VC; Yl.wl_cap = v.wl_cap;
        Yl.o_pos = v.pos; Yl.o_spd = v.spd; Yl.o_rpos = v.rpos; Yl.o_vid = v.vid; Yl.o_lead = v.lead; Yl.o_foll = v.foll;
        Yl.o_drv = v.drv; Yl.o_pj = v.pj; Yl.o_blk = v.blk; Yl.o_imeta = v.hot_end;
        Yl.o_dn = v.dn; Yl.o_npos = v.npos; Yl.o_nspd = v.nspd; Yl.o_nblk = v.nblk; Yl.o_xlist = v.xlist; Yl.o_cnt = v.end;
    }
    const Layout &Y = Yl;
    extern __shared__ __align__(16) unsigned char smem_block[];
    unsigned char *const smem = GMEM ? a.workspace + (size_t) blockIdx.x * (size_t) ((Y.smem_bytes + 255) & ~255) : smem_block;
    // GMEM with meta_shared (the usual case: the per-vehicle columns of a very large replica do not fit shared memory, its
    // per-drivable / per-signal arrays, scratch lists and header do): everything from o_cnt on lives in shared memory
    const bool hyb = GMEM && Y.meta_shared;
    unsigned char *const msm = hyb ? smem_block - Y.o_cnt : smem;
    const int tid = threadIdx.x;
    Ctx c;
    c.vb = smem;
    c.h = hyb ? (RepHeader *) (smem_block + (Y.smem_bytes - Y.o_cnt)) : (RepHeader *) smem;
    c.cnt = (u16 *) (msm + Y.o_cnt); c.head = (u16 *) (msm + Y.o_head); c.tail = (u16 *) (msm + Y.o_tail);
    c.wq = (u16 *) (msm + Y.o_wq);
    c.sraw = msm + Y.o_sraw; c.scur = msm + Y.o_scur; c.schg = msm + Y.o_schg; c.stop = (int *) (msm + Y.o_stop);
    c.rpos = (int *) (smem + Y.o_rpos); c.vid = (int *) (smem + Y.o_vid);
    c.lead = (u16 *) (smem + Y.o_lead); c.foll = (u16 *) (smem + Y.o_foll); c.pj = smem + Y.o_pj;
    c.dn = (u32 *) (smem + Y.o_dn); c.xlist = (u16 *) (smem + Y.o_xlist);
    c.clist = (u16 *) (smem + Y.o_drv);     // the image's u16 drivable column is expanded into dn[]; its room is reused
    c.leave = msm + Y.o_leave; c.ent = msm + Y.o_ent; c.fresh = msm + Y.o_fresh;
    c.mv_slot = (u16 *) (msm + Y.o_mvslot); c.mv_to = (u16 *) (msm + Y.o_mvto); c.mv_q = (int *) (msm + Y.o_mvq); c.mv_pj = msm + Y.o_mvpj;
    c.scan = (int *) (msm + Y.o_scan);
    c.avail = (u32 *) (msm + Y.o_avail);
    c.sp_rec = (int *) (msm + Y.o_spawn); c.sp_lane = c.sp_rec + 4 * S.n_spawn_lanes; c.sp_base = c.sp_lane + S.n_spawn_lanes;
    for (int s = tid; s < S.n_spawn_lanes; s += NT) c.sp_lane[s] = __ldg(S.spawn_lane + s);
    if (ONE_T || S.T <= SMEM_TEMPLATES) {
        double *ts = (double *) (msm + Y.o_tmpl);
        for (int k = tid; k < S.T * TD_STRIDE; k += NT) ts[k] = __ldg(S.tmpl + k);
        c.tmpl = ts;
    } else c.tmpl = S.tmpl;
#ifdef TSC_PHASE_TIMING
    c.pt = a.phase_cycles;
    c.pt_last = clock64();
#endif
    // async_stage 2 (default): the image travels as bulk asynchronous copies (TMA, 1-D) issued by one thread, one per column
    __shared__ __align__(8) unsigned long long stage_bar;
    const bool bulk = !GMEM && Y.async_stage == 2;
    unsigned stage_parity = 0;
    if (bulk) {
        if (tid == 0) { mbar_init(&stage_bar, 1); fence_proxy_async(); }
        __syncthreads();
    }

    for (int b = a.b0 + blockIdx.x; b < a.B; b += gridDim.x) {
        unsigned char *img = images + (size_t) b * Y.img_bytes;
        // the kinematics ping-pong between two buffers every tick: start from the ones that mirror the image
        c.par = 0;
        set_pingpong(Y, c);
        c.ellt = (int *) (img + Y.o_ellt);      // cold column: worked on in place
        // ---- stage the replica image into the working set ----
        if (bulk) {
            if (tid == 0) {      // how many slots are in use is in the image's header: every column is one bulk copy of just that much
                const int n0 = *(const volatile int *) img;      // RepHeader::n_slots
                const unsigned n8 = (n0 * 8 + 15) & ~15, n4 = (n0 * 4 + 15) & ~15, n2 = (n0 * 2 + 15) & ~15, n1 = (n0 + 15) & ~15;
                fence_proxy_async();      // whatever the block read or wrote here before, ahead of the copy engine's writes
                mbar_expect_tx(&stage_bar, (unsigned) sizeof(RepHeader) + (unsigned) Y.meta_bytes + 2 * n8 + 2 * n4 + 3 * n2 + n1);
                bulk_g2s(c.h, img, (unsigned) sizeof(RepHeader), &stage_bar);
                bulk_g2s(msm + Y.o_cnt, img + Y.o_imeta, (unsigned) Y.meta_bytes, &stage_bar);
                bulk_g2s(smem + Y.o_pos, img + Y.o_pos, n8, &stage_bar);
                bulk_g2s(smem + Y.o_spd, img + Y.o_spd, n8, &stage_bar);
                bulk_g2s(smem + Y.o_rpos, img + Y.o_rpos, n4, &stage_bar);
                bulk_g2s(smem + Y.o_vid, img + Y.o_vid, n4, &stage_bar);
                bulk_g2s(smem + Y.o_lead, img + Y.o_lead, n2, &stage_bar);
                bulk_g2s(smem + Y.o_foll, img + Y.o_foll, n2, &stage_bar);
                bulk_g2s(smem + Y.o_blk, img + Y.o_blk, n2, &stage_bar);
                bulk_g2s(smem + Y.o_pj, img + Y.o_pj, n1, &stage_bar);
            }
            {   // meanwhile: per-tick counters start from zero (the list surgery of every tick leaves them that way)
                int4 *z = (int4 *) (msm + Y.o_leave);
                const int nz = (Y.o_fresh - Y.o_leave + ((S.L + 15) & ~15)) / 16;      // leave, ent, fresh are adjacent
                for (int k = tid; k < nz; k += NT) z[k] = make_int4(0, 0, 0, 0);
            }
            mbar_wait(&stage_bar, stage_parity);
            stage_parity ^= 1u;
        } else {
            copy16(c.h, img, (int) sizeof(RepHeader), tid, NT);
            copy16(msm + Y.o_cnt, img + Y.o_imeta, Y.meta_bytes, tid, NT);
            __syncthreads();
        }
        const int n_in = c.h->n_slots;
        c.lso = S.lane_spawn_off + (size_t) c.h->flow_set * (S.L + 1);
        if (!bulk) {
            const int n8 = (n_in * 8 + 15) & ~15, n4 = (n_in * 4 + 15) & ~15, n2 = (n_in * 2 + 15) & ~15, n1 = (n_in + 15) & ~15;
            if (!GMEM && Y.async_stage) {
                copy16_async(smem + Y.o_pos, img + Y.o_pos, n8, tid, NT);
                copy16_async(smem + Y.o_spd, img + Y.o_spd, n8, tid, NT);
                copy16_async(smem + Y.o_rpos, img + Y.o_rpos, n4, tid, NT);
                copy16_async(smem + Y.o_vid, img + Y.o_vid, n4, tid, NT);
                copy16_async(smem + Y.o_lead, img + Y.o_lead, n2, tid, NT);
                copy16_async(smem + Y.o_foll, img + Y.o_foll, n2, tid, NT);
                copy16_async(smem + Y.o_blk, img + Y.o_blk, n2, tid, NT);
                copy16_async(smem + Y.o_pj, img + Y.o_pj, n1, tid, NT);
            } else {
                copy16(smem + Y.o_pos, img + Y.o_pos, n8, tid, NT);
                copy16(smem + Y.o_spd, img + Y.o_spd, n8, tid, NT);
                copy16(smem + Y.o_rpos, img + Y.o_rpos, n4, tid, NT);
                copy16(smem + Y.o_vid, img + Y.o_vid, n4, tid, NT);
                copy16(smem + Y.o_lead, img + Y.o_lead, n2, tid, NT);
                copy16(smem + Y.o_foll, img + Y.o_foll, n2, tid, NT);
                copy16(smem + Y.o_blk, img + Y.o_blk, n2, tid, NT);
                copy16(smem + Y.o_pj, img + Y.o_pj, n1, tid, NT);
            }
        }
        if (!bulk) {   // per-tick counters start from zero (the list surgery of every tick leaves them that way)
            int4 *z = (int4 *) (msm + Y.o_leave);
            const int nz = (Y.o_fresh - Y.o_leave + ((S.L + 15) & ~15)) / 16;      // leave, ent, fresh are adjacent
            for (int k = tid; k < nz; k += NT) z[k] = make_int4(0, 0, 0, 0);
        }
        if (tid == 0) { c.h->n_ent = 0; c.h->n_x = 0; c.h->n_h = 0; c.h->n_a = 0; c.h->n_new = n_in; }
        for (int s = tid; s < S.n_spawn_lanes; s += NT) {      // the head of every waiting buffer, in this replica's flow set
            const int l = __ldg(S.spawn_lane + s);
            const int b0 = __ldg(c.lso + l), b1 = __ldg(c.lso + l + 1);
            c.sp_base[s] = b0; c.sp_base[s + S.n_spawn_lanes] = b1;
            const int at = b0 + c.wq[s];
            ((int4 *) c.sp_rec)[s] = at < b1 ? __ldg(S.spawn_rec + at) : make_int4(-1, INT_MAX, 0, 0);
        }
        if (!GMEM && Y.async_stage == 1) copy16_async_wait();      // this thread's requests; the barrier publishes everybody's
        __syncthreads();
        // drivable | next drivable: the route table is read once per vehicle per launch, not once per tick
        {
            const u16 *drv16 = (const u16 *) (img + Y.o_drv);
            for (int i = tid; i < n_in; i += NT) {
                u32 nd = 0xFFFFu;
                if (c.vid[i] >= 0) nd = (u32) (__ldg(S.route_seq + c.rpos[i] + 1) & 0xFFFF);
                c.dn[i] = (u32) drv16[i] | (nd << 16);
            }
        }
        if (!GMEM && Y.prefetch_next) {   // this block's next replica: pull its image into L2 while this one is stepped
            const int nb = b + gridDim.x;
            if (nb < a.B) {
                const unsigned char *nimg = images + (size_t) nb * Y.img_bytes;
                for (int o = tid * 128; o < Y.img_bytes; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nimg + o));
            }
        }
        __syncthreads();
        pt_mark(c, PT_STAGE_IN);

        int *decided = (int *) c.nspd;       // free between ticks
        if (CTL && (a.apply_actions >= 4 || a.decide_only)) controller_decide<NT>(S, c, a, b, decided);
        if (!a.decide_only) apply_controller<NT>(S, c, a, b, decided);
        __syncthreads();
        // lane-link availability under the signals' current light phases: fixed for the whole launch
        for (int k0 = 0; k0 < S.K; k0 += NT) {      // uniform trip count: every lane takes part in the ballot
            const int k = k0 + tid;
            bool on = false;
            if (k < S.K) {
                int sb = __ldg(&S.llinfo[k].sigbit);
                int sg = sb & 0xFFFF;
                on = (__ldg(S.sig_phase_mask + sg * S.max_raw + c.sraw[sg]) >> (sb >> 16)) & 1u;
            }
            const u32 bits = __ballot_sync(0xffffffffu, on);
            if ((tid & 31) == 0 && k < S.K) c.avail[k >> 5] = bits;
        }
        bool frozen = __syncthreads_or(c.h->err != 0) != 0;
        pt_mark(c, PT_PROLOGUE);
        for (int t = 0; t < a.n_ticks; ++t) frozen = engine_tick<NT, ONE_T>(S, Y, c, frozen);
        if (a.do_retrieve) retrieve<NT>(S, Y, c, a, b, smem);
        pt_mark(c, PT_RETRIEVE);

        // ---- write the image back ----
        if (bulk && !a.decide_only && a.n_ticks > 0) {
            // the columns leave as bulk copies too: every thread orders its own shared-memory writes ahead of the copy
            // engine's reads, one thread issues; it waits until the engine has READ the working set before anything reuses it
            fence_proxy_async();
            __syncthreads();
            const int n = c.h->n_slots;
            if (tid == 0) {
                const unsigned n8 = (n * 8 + 15) & ~15, n4 = (n * 4 + 15) & ~15, n2 = (n * 2 + 15) & ~15, n1 = (n + 15) & ~15;
                bulk_s2g(img, c.h, (unsigned) sizeof(RepHeader));
                bulk_s2g(img + Y.o_imeta, msm + Y.o_cnt, (unsigned) Y.meta_bytes);
                bulk_s2g(img + Y.o_pos, c.pos, n8);      // whichever buffer holds the current state
                bulk_s2g(img + Y.o_spd, c.spd, n8);
                bulk_s2g(img + Y.o_blk, c.blk, n2);
                bulk_s2g(img + Y.o_rpos, smem + Y.o_rpos, n4);
                bulk_s2g(img + Y.o_vid, smem + Y.o_vid, n4);
                bulk_s2g(img + Y.o_lead, smem + Y.o_lead, n2);
                bulk_s2g(img + Y.o_foll, smem + Y.o_foll, n2);
                bulk_s2g(img + Y.o_pj, smem + Y.o_pj, n1);
                bulk_commit();
            }
            u32 *drv_pairs = (u32 *) (img + Y.o_drv);     // two u16 drivables per 32-bit store
            for (int i = tid; i < (n + 1) / 2; i += NT) {
                u32 lo = c.dn[2 * i] & 0xFFFFu;
                u32 hi = 2 * i + 1 < n ? (c.dn[2 * i + 1] & 0xFFFFu) : 0u;
                drv_pairs[i] = lo | (hi << 16);
            }
            if (tid == 0) bulk_wait_read();
        } else if (!a.decide_only && (a.n_ticks > 0 || a.apply_actions || a.set_raw_phase || a.init_program >= 0)) {
            __syncthreads();      // retrieve's scratch lives in the next-state buffers; nothing below reads them
            copy16(img, c.h, (int) sizeof(RepHeader), tid, NT);
            copy16(img + Y.o_imeta, msm + Y.o_cnt, Y.meta_bytes, tid, NT);
            const int n = c.h->n_slots;
            const int n8 = (n * 8 + 15) & ~15, n4 = (n * 4 + 15) & ~15, n2 = (n * 2 + 15) & ~15, n1 = (n + 15) & ~15;
            if (a.n_ticks > 0) {
                copy16(img + Y.o_pos, c.pos, n8, tid, NT);      // whichever buffer holds the current state
                copy16(img + Y.o_spd, c.spd, n8, tid, NT);
                copy16(img + Y.o_blk, c.blk, n2, tid, NT);
                copy16(img + Y.o_rpos, smem + Y.o_rpos, n4, tid, NT);
                copy16(img + Y.o_vid, smem + Y.o_vid, n4, tid, NT);
                copy16(img + Y.o_lead, smem + Y.o_lead, n2, tid, NT);
                copy16(img + Y.o_foll, smem + Y.o_foll, n2, tid, NT);
                copy16(img + Y.o_pj, smem + Y.o_pj, n1, tid, NT);
                u32 *drv_pairs = (u32 *) (img + Y.o_drv);     // two u16 drivables per 32-bit store
                for (int i = tid; i < (n + 1) / 2; i += NT) {
                    u32 lo = c.dn[2 * i] & 0xFFFFu;
                    u32 hi = 2 * i + 1 < n ? (c.dn[2 * i + 1] & 0xFFFFu) : 0u;
                    drv_pairs[i] = lo | (hi << 16);
                }
            }
        }
        __syncthreads();
        pt_mark(c, PT_STAGE_OUT);
    }
    if (bulk && tid == 0) bulk_wait_all();      // the last image's stores are complete before the block retires
}

// ----------------------------------------------------------------------------
// MetricsParser.mst: maximum spanning forest of a replica's density map (Prim, one block per replica)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tsc_mst_kernel(const double *dm, double *out, int A) {
    extern __shared__ __align__(16) unsigned char mst_smem[];
    double *key = (double *) mst_smem;            // [A] heaviest edge from the tree to the vertex (0 = none yet)
    int *parent = (int *) (key + A);              // [A]
    int *in_tree = parent + A;                    // [A]
    __shared__ double red_w[8];
    __shared__ int red_v[8];
    __shared__ int pick;
    const double *W = dm + (size_t) blockIdx.x * A * A;
    double *O = out + (size_t) blockIdx.x * A * A;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int p = tid; p < A * A; p += 256) O[p] = 0.0;
    for (int v = tid; v < A; v += 256) { key[v] = 0.0; parent[v] = -1; in_tree[v] = 0; }
    __syncthreads();
    for (int it = 0; it < A; ++it) {
        // the vertex outside the tree with the heaviest edge into it; none reachable: the lowest outsider (a new component)
        double bw = -1.0;
        int bv = INT_MAX;
        for (int v = tid; v < A; v += 256)
            if (!in_tree[v] && (key[v] > bw || (key[v] == bw && v < bv))) { bw = key[v]; bv = v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ow = __shfl_xor_sync(0xffffffffu, bw, o);
            const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
            if (ow > bw || (ow == bw && ov < bv)) { bw = ow; bv = ov; }
        }
        if (lane == 0) { red_w[w] = bw; red_v[w] = bv; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < 8; ++k)
                if (red_w[k] > bw || (red_w[k] == bw && red_v[k] < bv)) { bw = red_w[k]; bv = red_v[k]; }
            pick = bv;
            in_tree[bv] = 1;
            if (bw > 0.0) {
                const int p = parent[bv];
                O[(size_t) min(p, bv) * A + max(p, bv)] = -bw;
            }
        }
        __syncthreads();
        const int u = pick;
        for (int v = tid; v < A; v += 256) {
            const double wt = W[(size_t) u * A + v];
            if (!in_tree[v] && wt > key[v]) { key[v] = wt; parent[v] = u; }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------
// Host side: handle, tables, C ABI
// ----------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (x);                                                                        \
        if (e__ != cudaSuccess)                                                                       \
            return fail(e__ == cudaErrorMemoryAllocation ? TSC_ENOMEM : TSC_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e__)); \
    } while (0)

// Kernel variants, picked from how many replica working sets fit an SM's shared memory: four 192-thread
// blocks per SM (80 registers) -- the bench workload --, three or two 256-thread blocks (80 / 128 registers),
// one 512- or 1024-thread block, or one 1024-thread block over a global-memory workspace for replicas that do
// not fit shared memory at all.  Scenarios with several vehicle templates run the generic (per-vehicle
// template look-up) builds of the 256 x 2, 512 and global-memory variants.
typedef void (*step_kernel_t)(const DevScn, const Layout, unsigned char *, const StepArgs);
// capacities (slots, after tsc_create's rounding) with a fixed-capacity build of the variant they fit: the bench workloads
#define VC_HANGZHOU 672      // vehicle_capacity 640: 256 x 4
#define VC_JINAN 1184        // vehicle_capacity 1150: 256 x 3
#define VC_MANHATTAN 1568    // vehicle_capacity 1530: 384 x 2
static step_kernel_t kernel_for(int nt, int minb, bool ctl, bool one_t, bool gmem, int vc) {
    if (!gmem && one_t && nt == 256 && minb >= 4 && vc == VC_HANGZHOU)
        return ctl ? tsc_step_kernel<256, 4, true, false, true, VC_HANGZHOU> : tsc_step_kernel<256, 4, false, false, true, VC_HANGZHOU>;
    if (!gmem && one_t && nt == 256 && minb == 3 && vc == VC_JINAN)
        return ctl ? tsc_step_kernel<256, 3, true, false, true, VC_JINAN> : tsc_step_kernel<256, 3, false, false, true, VC_JINAN>;
    if (!gmem && one_t && nt == 384 && vc == VC_MANHATTAN)
        return ctl ? tsc_step_kernel<384, 2, true, false, true, VC_MANHATTAN> : tsc_step_kernel<384, 2, false, false, true, VC_MANHATTAN>;
    if (gmem && one_t) return ctl ? tsc_step_kernel<1024, 1, true, true, true> : tsc_step_kernel<1024, 1, false, true, true>;
    if (gmem) return tsc_step_kernel<1024, 1, true, true, false>;
    if (!one_t) return nt >= 512 ? tsc_step_kernel<512, 1, true, false, false> : tsc_step_kernel<256, 2, true, false, false>;
    if (nt == 1024) return tsc_step_kernel<1024, 1, true, false, true>;      // 32 warps at 64 registers
    if (nt == 512) return tsc_step_kernel<512, 1, true, false, true>;
    if (nt == 160) return ctl ? tsc_step_kernel<160, 5, true, false, true> : tsc_step_kernel<160, 5, false, false, true>;
    if (nt == 192 && minb >= 5) return ctl ? tsc_step_kernel<192, 5, true, false, true> : tsc_step_kernel<192, 5, false, false, true>;
    if (nt == 192) return ctl ? tsc_step_kernel<192, 4, true, false, true> : tsc_step_kernel<192, 4, false, false, true>;
    if (nt == 384) return ctl ? tsc_step_kernel<384, 2, true, false, true> : tsc_step_kernel<384, 2, false, false, true>;
    if (nt == 256 && minb >= 4) return ctl ? tsc_step_kernel<256, 4, true, false, true> : tsc_step_kernel<256, 4, false, false, true>;
    if (ctl) return minb >= 3 ? tsc_step_kernel<256, 3, true, false, true> : tsc_step_kernel<256, 2, true, false, true>;
    return minb >= 3 ? tsc_step_kernel<256, 3, false, false, true> : tsc_step_kernel<256, 2, false, false, true>;
}

#define MAX_HOST_CHUNKS 16

struct tsc_engine {
    int device = 0, B = 0;
    DevScn S{};
    Layout Y{};
    std::vector<void *> dev_allocs;
    unsigned char *images = nullptr;
    int *d_actions = nullptr;
    float *d_obs = nullptr, *d_reward = nullptr, *d_rg = nullptr;
    u8 *d_mask = nullptr;
    int32_t *h_actions = nullptr;      // pinned staging for the *_host path
    float *h_obs = nullptr, *h_reward = nullptr, *h_rg = nullptr;
    u8 *h_mask = nullptr;
    int grid = 0, grid_ctl = 0, regs = 0, nt = 256, minb = 2;
    step_kernel_t kern = nullptr, kern_ctl = nullptr;   // plain variant / with the rule-based controllers
    int64_t launches = 0;
    unsigned long long *d_phase_cycles = nullptr;   // debug phase timing buffer (tsc_debug_timing)
    unsigned char *workspace = nullptr;             // GMEM variant: grid working sets in global memory
    bool gmem = false;
    int dyn_smem = 0;                               // dynamic shared memory per block of the chosen variant
    cudaStream_t host_compute = nullptr, host_compute2 = nullptr, host_copy = nullptr;   // tsc_env_step_host: step chunk k+1 while chunk k is copied out
    int host_streams = 1;               // compute streams the chunks alternate on (TSC_B200_HOST_STREAMS=2: measured slower, 1.58 vs 1.53 ms per B=4096 step)
    cudaEvent_t host_ev_actions = nullptr;
    cudaEvent_t host_ev[1 + MAX_HOST_CHUNKS] = {};
    int host_chunks = MAX_HOST_CHUNKS;   // upper bound on chunks per host step (TSC_B200_HOST_CHUNKS)
    int host_lead = 0;                  // TSC_B200_HOST_LEAD=1 / =N: a short first chunk (one replica per SM / N replicas) so that the copies start early -- measured slower (1.58 vs 1.39-1.50 ms)
    bool host_zero_copy = false;        // TSC_B200_HOST_ZERO_COPY=1: kernel stores straight into mapped page-locked buffers
    std::vector<unsigned char> init_image;   // host copy of the tick-0 image
    std::vector<int> h_flow_set;             // [B] flow set every replica runs (applied at reset)
    int *d_flow_set = nullptr;
    struct HostPath *hp = nullptr;           // registered end-to-end path (tsc_host_register)
    bool fixed_capacity = false;             // the step kernel is a fixed-capacity build (per-vehicle columns at compile-time offsets)
    int host_threads_req = 0;                // tsc_host_threads: workers of the registered path (0 = automatic)
    std::vector<u32> h_obs_code;             // host copies of the observation recipe (tsc_host_register fills the static columns)
    std::vector<float> h_obs_static;
    std::vector<int> h_pk_dst;               // [n_in_total] float offset of the lane's n_queued element inside a replica's observation block, -1 = not shown
    std::vector<int> h_sig_n_phases;
    // host copies needed by snapshot/load
    std::vector<int> h_route_seq, h_veh_seq_start;
    int n_spawn_lanes = 0;
    std::vector<int> h_spawn_lane;
    std::vector<u8> h_is_spawn;
};


// ----------------------------------------------------------------------------
// Registered end-to-end path (tsc_host_register / tsc_env_step_registered)
// ----------------------------------------------------------------------------
// The launch stores one compact packet per replica into mapped page-locked host memory and raises the
// replica's flag behind a system-scope fence (retrieve()).  Host worker threads follow the flags while the
// launch is still running and finish the caller's rows: the packets ping-pong between two buffers, so the
// previous step's packet is the shadow copy that tells which values changed.
struct HostPath {
    tsc_engine *E = nullptr;
    float *obs = nullptr, *reward = nullptr, *rg = nullptr;
    u8 *mask = nullptr;
    unsigned char *pk[2] = {nullptr, nullptr};       // host addresses
    unsigned char *pk_dev[2] = {nullptr, nullptr};   // the same buffers as the device sees them
    u32 *flags = nullptr, *flags_dev = nullptr;
    u32 seq = 0;
    int cur = 0;
    int nthreads = 1;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    u32 job_seq = 0;          // guarded by mu: the step the workers should process
    bool quit = false;
    bool pending = false;     // a step begun and not yet waited for
    // work sharing: groups of HOST_GROUP replicas are claimed one at a time by whoever has time (the workers, and the caller
    // while it waits); both words carry the step's sequence number in their upper half, so a straggler of the previous step
    // can neither claim nor count anything of this one
    std::atomic<uint64_t> claim{0};           // (seq << 32) | next group to hand out
    std::atomic<uint64_t> groups_done{0};     // (seq << 32) | groups finished
    std::atomic<int> failed{0};
    uint64_t mask_lut[256];   // byte k of entry x = bit k of x
};

static const int HOST_GROUP = 16;      // consecutive replicas a worker takes at a time

// Finish replica b's rows in the caller's buffers from its packet (new) against the previous one (old).
static void host_finish_replica(HostPath *H, int b) {
    tsc_engine *E = H->E;
    const DevScn &S = E->S;
    const int A = S.A, row = S.state_dim, pkb = S.pk_bytes;
    const unsigned char *np = H->pk[H->cur] + (size_t) b * pkb, *op = H->pk[H->cur ^ 1] + (size_t) b * pkb;
    if (H->obs) {
        float *ob = H->obs + (size_t) b * A * row;
        const int *dst = E->h_pk_dst.data();
        const int n = S.n_in_total;
        if (S.pk_mode) {
            const u32 *nw = (const u32 *) np, *ow = (const u32 *) op;
            // pass 1: which lanes changed (most keep their values from one step to the next: four at a time); the row
            // lines they live in are requested for writing all at once -- the arrays are far larger than the caches, and one
            // miss after the other was most of this function's time.  pass 2: the writes.
            int changed[512];
            int nc = 0;
            for (int i0 = 0; i0 < n; i0 += 4) {
                const int i1 = i0 + 4 < n ? i0 + 4 : n;
                if (i1 - i0 == 4 && memcmp(nw + i0, ow + i0, 16) == 0) continue;
                for (int i = i0; i < i1; ++i) {
                    if (nw[i] != ow[i] && dst[i] >= 0) {
                        if (nc == 512) {      // (more than 512 changed lanes in one replica: flush)
                            for (int k = 0; k < nc; ++k) {
                                const u32 v = nw[changed[k]];
                                float *d = ob + dst[changed[k]];
                                d[0] = (float) (v & 255u); d[1] = (float) ((v >> 8) & 255u); d[2] = (float) ((v >> 16) & 255u);
                            }
                            nc = 0;
                        }
                        __builtin_prefetch(ob + dst[i], 1);
                        changed[nc++] = i;
                    }
                }
            }
            for (int k = 0; k < nc; ++k) {
                const u32 v = nw[changed[k]];
                float *d = ob + dst[changed[k]];
                d[0] = (float) (v & 255u); d[1] = (float) ((v >> 8) & 255u); d[2] = (float) ((v >> 16) & 255u);
            }
        } else {
            const u32 *nw = (const u32 *) np, *ow = (const u32 *) op;      // compared as bit patterns
            for (int i = 0; i < n; ++i) {
                if ((nw[3 * i] != ow[3 * i] || nw[3 * i + 1] != ow[3 * i + 1] || nw[3 * i + 2] != ow[3 * i + 2]) && dst[i] >= 0)
                    memcpy(ob + dst[i], nw + 3 * i, 12);
            }
        }
        const u8 *nph = np + S.pk_o_phase, *oph = op + S.pk_o_phase;
        const int hot0 = S.max_lanes_per_signal * 12;
        for (int sg = 0; sg < A; ++sg) {
            if (nph[sg] != oph[sg]) {      // phase one-hot (observations.py:322-324): move the 1
                float *d = ob + (size_t) sg * row + hot0;
                d[oph[sg]] = 0.0f; d[nph[sg]] = 1.0f;
            }
        }
    }
    if (H->reward) memcpy(H->reward + (size_t) b * A, np + S.pk_o_reward, (size_t) A * 4);
    if (H->rg) H->rg[b] = *(const float *) (np + S.pk_o_rg);
    if (H->mask) {
        const u32 *bits = (const u32 *) (np + S.pk_o_mask), *obits = (const u32 *) (op + S.pk_o_mask);
        const int na = S.n_actions;
        u8 *m = H->mask + (size_t) b * A * na;
        const bool first = H->seq <= 2;      // (the two packet buffers start zeroed, the caller's mask array does not)
        for (int sg = 0; sg < A; ++sg, m += na) {
            const u32 x = bits[sg];
            if (!first && x == obits[sg]) continue;
            int p = 0;
            for (; p + 8 <= na; p += 8) memcpy(m + p, &H->mask_lut[(x >> p) & 255u], 8);
            for (; p < na; ++p) m[p] = (u8) ((x >> p) & 1u);
        }
    }
}

static void host_work(HostPath *H, u32 seq) {
    const int B = H->E->B;
    const u32 ngroups = (u32) ((B + HOST_GROUP - 1) / HOST_GROUP);
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        uint64_t v = H->claim.load(std::memory_order_acquire);
        for (;;) {
            if ((u32) (v >> 32) != seq || (u32) v >= ngroups) return;      // another step's turn, or nothing left to hand out
            if (H->claim.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) break;
        }
        const int g = (int) (u32) v;
        const int b1 = std::min(B, (g + 1) * HOST_GROUP);
        for (int b = g * HOST_GROUP; b < b1; ++b) {
            unsigned spins = 0;
            while (__atomic_load_n(&H->flags[b], __ATOMIC_ACQUIRE) != seq) {
                // a short spin covers the gaps inside a running launch; a launch that has not started yet (the other handle's is
                // still running, or the stream is busy) is waited for asleep, so that the core goes to the policy thread or to the
                // workers of the handle whose packets ARE arriving
                if (++spins < 512) __builtin_ia32_pause();
                else { struct timespec ts = {0, 15000}; nanosleep(&ts, nullptr); }
                if ((spins & 0x3FF) == 0) {
                    if (H->failed.load(std::memory_order_relaxed)) return;
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) { H->failed.store(2); return; }
                }
            }
            host_finish_replica(H, b);
        }
        H->groups_done.fetch_add(1, std::memory_order_release);
    }
}

static void host_worker_main(HostPath *H) {
    prctl(PR_SET_TIMERSLACK, 2000UL, 0, 0, 0);      // the 15 us naps above are not to be rounded up to the default 50 us slack
    u32 seen = 0;
    for (;;) {
        u32 seq;
        {
            std::unique_lock<std::mutex> lk(H->mu);
            H->cv.wait(lk, [&] { return H->quit || H->job_seq != seen; });
            if (H->quit) return;
            seq = seen = H->job_seq;
        }
        host_work(H, seq);
    }
}

static int host_thread_count() {
    if (const char *env = getenv("TSC_B200_HOST_THREADS")) { int v = atoi(env); if (v >= 1 && v <= 64) return v; }
    int cpus = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
    if (cpus <= 0) cpus = (int) std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char *env = getenv("LOCAL_WORLD_SIZE")) { int v = atoi(env); if (v >= 1) ranks = v; }
    int n = cpus / ranks;
    return n < 1 ? 1 : (n > 8 ? 8 : n);
}

template <typename Tp>
static int upload(tsc_engine *E, const Tp *host, size_t n, const Tp **dev) {
    void *p = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(Tp);
    CUDA_TRY(cudaMalloc(&p, bytes));
    E->dev_allocs.push_back(p);
    if (n) CUDA_TRY(cudaMemcpy(p, host, n * sizeof(Tp), cudaMemcpyHostToDevice));
    *dev = (const Tp *) p;
    return 0;
}

static int align16(int x) { return (x + 15) & ~15; }

static void build_layout(Layout &Y, const DevScn &S, int Vcap, int n_warps) {
    Y.Vcap = Vcap;
    // vehicles changing drivable in one tick (a lane hands over at most one or two per tick; the shipped
    // workloads stay below a twentieth of the running vehicles): an eighth of the slots, overflow is
    // reported (ERR_ENT_OVERFLOW)
    Y.ent_cap = Vcap / 8 < 64 ? 64 : (Vcap / 8 > 8192 ? 8192 : Vcap / 8);
    // between ticks the next-state kinematics pair (16 bytes per slot, contiguous) is the scratch of retrieve /
    // the controllers: lane sums, per-signal terms, then the position-matrix windows or the host packet
    const int tailb = S.obs_type == TSC_OBS_POSITION_MATRIX ? std::max(S.n_in_total * S.visibility * 8, S.pk_bytes) : S.pk_bytes;
    const int scratch = 24 * S.L + 16 * S.A + 4 * ((S.L + 1) & ~1) + tailb + 64;
    int Vlay = Vcap;
    if (16 * Vlay < scratch) Vlay = ((scratch + 15) / 16 + 7) & ~7;
    Y.Vlay = Vlay;
    // ---- per-vehicle columns (image: the hot ones; working set: all of them)
    const VOff v = v_offsets(Vcap, Vlay, n_warps);
    Y.o_pos = v.pos; Y.o_spd = v.spd; Y.o_rpos = v.rpos; Y.o_vid = v.vid; Y.o_lead = v.lead; Y.o_foll = v.foll;
    Y.o_drv = v.drv; Y.o_pj = v.pj; Y.o_blk = v.blk;
    Y.o_dn = v.dn; Y.o_npos = v.npos; Y.o_nspd = v.nspd; Y.o_nblk = v.nblk; Y.o_xlist = v.xlist; Y.wl_cap = v.wl_cap;
    // ---- per-drivable / per-signal block: behind the per-vehicle columns of the working set ...
    int o = v.end;
    Y.o_cnt = o; o = align16(o + 2 * (S.D + 2));
    Y.o_head = o; o = align16(o + 2 * (S.D + 2));
    Y.o_tail = o; o = align16(o + 2 * (S.D + 2));
    Y.o_wq = o; o = align16(o + 2 * (S.n_spawn_lanes + 1));
    Y.o_sraw = o; o = align16(o + S.A);
    Y.o_scur = o; o = align16(o + S.A);
    Y.o_schg = o; o = align16(o + S.A);
    Y.o_stop = o; o = align16(o + 4 * S.A);
    Y.o_meta_end = o;
    Y.meta_bytes = Y.o_meta_end - Y.o_cnt;
    // ... and behind the hot columns of the image, followed by the cold column (enterLaneLinkTime: read by canPass
    // tie-breaks, written when a vehicle enters a lane-link), which stays in the image and is worked on in place
    Y.o_imeta = v.hot_end;
    Y.o_ellt = Y.o_imeta + Y.meta_bytes;
    Y.img_bytes = align16(Y.o_ellt + 4 * Vcap);
    // ---- working-set-only scratch
    Y.o_leave = o; o = align16(o + S.D + 4);      // leave | ent | fresh adjacent: zeroed together at stage-in
    Y.o_ent = o; o = align16(o + S.D + 4);
    Y.o_fresh = o; o = align16(o + S.L);
    Y.o_mvslot = o; o = align16(o + 2 * Y.ent_cap);
    Y.o_mvto = o; o = align16(o + 2 * Y.ent_cap);
    Y.o_mvq = o; o = align16(o + 4 * Y.ent_cap);
    Y.o_mvpj = o; o = align16(o + Y.ent_cap);
    Y.o_scan = o; o = align16(o + 4 * 64);
    Y.o_avail = o; o = align16(o + 4 * ((S.K + 31) / 32 + 1));
    Y.o_tmpl = o; o = align16(o + 8 * TD_STRIDE * (S.T < SMEM_TEMPLATES ? S.T : SMEM_TEMPLATES));
    Y.o_spawn = o; o = align16(o + 28 * (S.n_spawn_lanes + 1));      // 16-byte records first, then the lanes and the record ranges
    Y.smem_bytes = o;
}

// byte offset IN THE IMAGE of a per-drivable / per-signal array given by its working-set offset
static inline int img_meta(const Layout &Y, int ws_off) { return ws_off - Y.o_cnt + Y.o_imeta; }

extern "C" {

int tsc_abi_version(void) { return TSC_ABI_VERSION; }
const char *tsc_last_error(void) { return g_err.c_str(); }

static int create_body(tsc_engine *E, const tsc_scenario_t *s, int32_t n_replicas, int32_t device, int32_t vehicle_capacity);

int tsc_create(const tsc_scenario_t *s, int32_t n_replicas, int32_t device, int32_t vehicle_capacity, tsc_handle *out) {
    if (!s || !out) return fail(TSC_EINVAL, "null argument");
    *out = nullptr;
    if (s->abi_version != TSC_ABI_VERSION) return fail(TSC_EINVAL, "scenario abi_version %d != %d", s->abi_version, TSC_ABI_VERSION);
    if (n_replicas <= 0) return fail(TSC_EINVAL, "n_replicas must be positive");
    if (s->interval != 1.0) return fail(TSC_EINVAL, "interval must be 1.0");
    if (s->visibility > 16 || s->visibility < 1) return fail(TSC_EINVAL, "visibility must be in 1..16");
    if (s->max_phases > 32) return fail(TSC_EINVAL, "more than 32 phases per signal");
    if (s->n_lanes + s->n_lanelinks >= 65535) return fail(TSC_EINVAL, "too many drivables for one replica block");
    if (!s->ctl_off || s->ctl_off[s->n_signals * s->max_phases] != s->n_ctl_total) return fail(TSC_EINVAL, "ctl_off does not end at n_ctl_total");
    for (int e = 0; e < s->n_ctl_total; ++e)
        if (s->ctl_in_lane[e] < 0 || s->ctl_in_lane[e] >= s->n_lanes || s->ctl_out_lane[e] >= s->n_lanes)
            return fail(TSC_EINVAL, "controller table entry %d: bad lane index", e);
    if (!s->dm_off || s->dm_off[(size_t) s->n_signals * s->n_signals] != s->n_dm_total) return fail(TSC_EINVAL, "dm_off does not end at n_dm_total");
    for (int e = 0; e < s->n_dm_total; ++e)
        if (s->dm_lane[e] < 0 || s->dm_lane[e] >= s->n_lanes) return fail(TSC_EINVAL, "density-map table entry %d: bad lane index", e);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(TSC_EINVAL, "device %d not available (%d devices)", device, ndev);
    if (s->n_flow_sets < 1) return fail(TSC_EINVAL, "n_flow_sets must be at least 1");
    for (int a = 0; a < s->n_signals; ++a)
        if (s->sig_n_raw_phases[a] < 1 || s->sig_n_raw_phases[a] > s->max_raw_phases || s->sig_n_phases[a] < 1 || s->sig_n_phases[a] > s->max_phases)
            return fail(TSC_EINVAL, "signal %d: phase counts out of range", a);
    CUDA_TRY(cudaSetDevice(device));
    auto *E = new tsc_engine();
    E->device = device; E->B = n_replicas;
    // every failure below releases the handle and whatever it already owns (device tables, pinned buffers, streams)
    const int rc = create_body(E, s, n_replicas, device, vehicle_capacity);
    if (rc) { tsc_destroy(E); return rc; }
    *out = E;
    return 0;
}

}  // extern "C"

static int create_body(tsc_engine *E, const tsc_scenario_t *s, int32_t n_replicas, int32_t device, int32_t vehicle_capacity) {
    DevScn &S = E->S;
    S.L = s->n_lanes; S.K = s->n_lanelinks; S.D = S.L + S.K; S.A = s->n_signals; S.N = s->n_vehicles; S.T = s->n_templates;
    S.F = s->n_flow_sets;
    S.horizon = s->horizon_ticks; S.max_raw = s->max_raw_phases; S.P = s->max_phases;
    S.n_in_total = s->n_in_total; S.n_out_total = s->n_out_total;
    const int L = S.L, K = S.K, D = S.D, A = S.A, N = S.N;
    int rc = 0;
#define UP(field, n) if ((rc = upload(E, s->field, (size_t) (n), &S.field))) return rc;
    UP(drv_length, D) UP(drv_max_speed, D) UP(lane_ll_off, L + 1) UP(lane_ll, s->lane_ll_off[L])
    UP(lane_spawn_off, (size_t) S.F * (L + 1)) UP(lane_spawn_vid, N)
    if ((rc = upload(E, s->sig_n_raw_phases, (size_t) A, &S.sig_n_raw))) return rc;
    for (int f = 0; f < S.F; ++f) {
        const int *row = s->lane_spawn_off + (size_t) f * (L + 1);
        if (row[0] != (f ? s->lane_spawn_off[(size_t) f * (L + 1) - 1] : 0) || row[L] > N) return fail(TSC_EINVAL, "lane_spawn_off row %d does not continue the previous one", f);
        for (int l = 0; l < L; ++l) if (row[l + 1] < row[l]) return fail(TSC_EINVAL, "lane_spawn_off row %d is not ascending", f);
    }
    UP(ll_start_lane, K) UP(ll_end_lane, K) UP(ll_signal, K) UP(ll_roadlink, K) UP(ll_type, K) UP(ll_cross_off, K + 1)
    UP(xr_dist, s->n_cross_entries) UP(xr_foe_ll, s->n_cross_entries) UP(xr_foe_dist, s->n_cross_entries)
    UP(sig_phase_mask, A * s->max_raw_phases) UP(route_seq, s->n_route_seq)
    UP(veh_tick, N) UP(veh_seq_start, N) UP(veh_tmpl, N) UP(veh_priority, N)
    UP(lane_pytsc_length, L) UP(lane_feat, L * 9) UP(sig_in_off, A + 1) UP(sig_in_lane, s->n_in_total)
    UP(sig_out_off, A + 1) UP(sig_out_lane, s->n_out_total) UP(sig_n_phases, A) UP(sig_phase_raw, A * s->max_phases)
    UP(sig_phase_green, A * s->max_phases) UP(sig_min_time, A * s->max_phases) UP(sig_max_time, A * s->max_phases)
    UP(nbr_off, A + 1) UP(nbr_idx, s->n_nbr_total) UP(nbr_weight, s->n_nbr_total)
    UP(ctl_off, A * s->max_phases + 1) UP(ctl_in_lane, s->n_ctl_total) UP(ctl_out_lane, s->n_ctl_total)
    UP(dm_off, (size_t) A * A + 1) UP(dm_lane, s->n_dm_total) UP(dm_adjacency, (size_t) A * A)
#undef UP
    {   // pytsc's lane length in vehicle cells (retriever.py:73), the same IEEE division the kernel used to repeat per lane
        std::vector<double> cells(L > 0 ? L : 1, 1.0);
        for (int l = 0; l < L; ++l) {
            cells[l] = s->lane_pytsc_length[l] / s->veh_size_min_gap;
            if (!(cells[l] > 0.0)) { return fail(TSC_EINVAL, "lane %d: pytsc length / veh_size_min_gap must be positive", l); }
        }
        if ((rc = upload(E, cells.data(), (size_t) L, &S.lane_cells))) return rc;
    }
    for (int k = 0; k < D; ++k)
        if (!(s->drv_max_speed[k] > 0.0)) { return fail(TSC_EINVAL, "drivable %d: max speed must be positive", k); }
    S.reward_type = s->reward_type; S.obs_type = s->obs_type; S.action_space = s->action_space; S.round_robin = s->round_robin;
    S.visibility = s->visibility; S.yellow_time = s->yellow_time; S.obs_dim = s->obs_dim; S.state_dim = s->state_dim;
    S.n_actions = s->n_actions; S.reference_exact = s->reference_exact; S.max_lanes_per_signal = s->max_lanes_per_signal;
    S.max_obs_phases = s->max_obs_phases; S.v_size = s->veh_size_min_gap; S.flick = s->flickering_coef; S.interval = s->interval;
    {   // lane-feature rows (observations.py:305-329): per incoming lane [9 static, n_queued, occupancy, mean_speed],
        // -1 padding up to max_lanes_per_signal lanes, then the phase one-hot over max_obs_phases entries
        const int per = 12, ML = S.max_lanes_per_signal, MP = S.max_obs_phases, row = ML * per + MP;
        if (row != S.state_dim) { return fail(TSC_EINVAL, "state_dim %d != %d * 12 + %d", S.state_dim, ML, MP); }
        if (S.obs_type == TSC_OBS_LANE_FEATURES && S.obs_dim != row) { return fail(TSC_EINVAL, "obs_dim %d != state_dim %d", S.obs_dim, row); }
        if (A > 0xFFFFF || MP > 256) { return fail(TSC_EINVAL, "too many signals / phases for the observation recipe"); }
        std::vector<u32> code((size_t) A * row > 0 ? (size_t) A * row : 1, 0u);
        std::vector<float> sval(code.size(), 0.0f);
        for (int sg = 0; sg < A; ++sg) {
            const int i0 = s->sig_in_off[sg], nin = s->sig_in_off[sg + 1] - i0;
            const bool tr = S.reference_exact && nin < ML;
            for (int k = 0; k < row; ++k) {
                const size_t at = (size_t) sg * row + k;
                if (k < ML * per) {
                    const int e = k / per, f = k - e * per;
                    if (e >= nin) { sval[at] = -1.0f; continue; }
                    const int l = s->sig_in_lane[i0 + e];
                    if (f < 9) { double x = s->lane_feat[(size_t) l * 9 + f]; sval[at] = (float) (tr ? trunc(x) : x); }
                    else code[at] = (u32) (f - 8) | (tr ? 8u : 0u) | ((u32) l << 4);
                } else {
                    const int ph = k - ML * per;
                    if (ph < s->sig_n_phases[sg]) code[at] = 4u | ((u32) (sg << 8 | ph) << 4);
                }
            }
        }
        if ((rc = upload(E, code.data(), code.size(), &S.obs_code))) return rc;
        if ((rc = upload(E, sval.data(), sval.size(), &S.obs_static))) return rc;
        E->h_obs_code = code; E->h_obs_static = sval;
        E->h_sig_n_phases.assign(s->sig_n_phases, s->sig_n_phases + A);
        // host packet of the registered end-to-end path: per incoming lane (observation-row order) the lane and whether
        // its row truncates; one u32 per lane when every shown value is a small integer, three floats otherwise
        std::vector<u32> pkl(s->n_in_total > 0 ? s->n_in_total : 1, 0u);
        E->h_pk_dst.assign(s->n_in_total > 0 ? s->n_in_total : 1, -1);
        double min_len = 1e300, max_speed = 0.0;
        for (int t = 0; t < s->n_templates; ++t) {
            min_len = std::min(min_len, s->tmpl[(size_t) t * TSC_T_STRIDE + TSC_T_LEN]);
            max_speed = std::max(max_speed, s->tmpl[(size_t) t * TSC_T_STRIDE + TSC_T_MAX_SPEED]);
        }
        bool small_ints = S.reference_exact != 0 && max_speed < 255.0 && min_len > 0.0;
        for (int sg = 0; sg < A; ++sg) {
            const int i0 = s->sig_in_off[sg], nin = s->sig_in_off[sg + 1] - i0;
            const bool tr = S.reference_exact && nin < ML;
            small_ints = small_ints && tr;
            for (int e = 0; e < nin; ++e) {
                const int l = s->sig_in_lane[i0 + e];
                pkl[i0 + e] = (u32) l | (tr ? 0x80000000u : 0u);
                if (e < ML) E->h_pk_dst[i0 + e] = sg * row + e * per + 9;
                // vehicles on the lane < 255, occupancy = n / cells <= n
                small_ints = small_ints && s->drv_length[l] / min_len + 2.0 < 255.0 && s->lane_pytsc_length[l] / s->veh_size_min_gap >= 1.0;
            }
        }
        if ((rc = upload(E, pkl.data(), pkl.size(), &S.pk_lane))) return rc;
        std::vector<int> insig(s->n_in_total > 0 ? s->n_in_total : 1, 0);
        for (int sg = 0; sg < A; ++sg)
            for (int e = s->sig_in_off[sg]; e < s->sig_in_off[sg + 1]; ++e) insig[e] = sg;
        if ((rc = upload(E, insig.data(), insig.size(), &S.in_sig))) return rc;
        S.pk_mode = small_ints ? 1 : 0;
        int o = align16((small_ints ? 4 : 12) * s->n_in_total);
        S.pk_o_phase = o; o = align16(o + A);
        S.pk_o_reward = o; o = align16(o + 4 * A);
        S.pk_o_mask = o; o = align16(o + 4 * A);
        S.pk_o_rg = o; o = align16(o + 4);
        S.pk_bytes = o;
    }
    {   // vehicle templates, extended with the per-template constants of the car-following law
        std::vector<double> td((size_t) s->n_templates * TD_STRIDE, 0.0);
        for (int t = 0; t < s->n_templates; ++t) {
            const double *src = s->tmpl + (size_t) t * TSC_T_STRIDE;
            double *dst = td.data() + (size_t) t * TD_STRIDE;
            for (int k = 0; k < TSC_T_STRIDE; ++k) dst[k] = src[k];
            if (!(src[TSC_T_MAX_NEG_ACC] > 0.0) || !(src[TSC_T_MAX_NEG_ACC] < 1e300)) { return fail(TSC_EINVAL, "template %d: maxNegAcc must be positive and finite", t); }
            const double a = 0.5 / src[TSC_T_MAX_NEG_ACC];
            dst[TD_A] = a;
            dst[TD_HALF_OVER_A] = 0.5 / a;
            dst[TD_HEADWAY_DEN] = src[TSC_T_HEADWAY] + s->interval / 2;
            {   // stop_before_speed with v = 0, operation for operation (volatile: no contraction, no folding surprises)
                volatile double v0 = 0.0, dt = 1.0;
                volatile double nxt = v0 + src[TSC_T_USUAL_POS_ACC] * dt;
                volatile double first = (v0 + nxt) * dt / 2, sq = nxt * nxt, q1 = sq / src[TSC_T_USUAL_NEG_ACC], second = q1 / 2;
                dst[TD_BRAKE0] = first + second;
            }
            const double approach = src[TSC_T_MAX_SPEED] * src[TSC_T_MAX_SPEED] / src[TSC_T_USUAL_NEG_ACC] / 2 + src[TSC_T_MAX_SPEED] * s->interval * 2;
            if (src[TSC_T_APPROACH_DIST] != approach) { return fail(TSC_EINVAL, "template %d: TSC_T_APPROACH_DIST is not maxSpeed^2/usualNegAcc/2 + 2 maxSpeed interval", t); }
        }
        if ((rc = upload(E, td.data(), td.size(), &S.tmpl))) return rc;
    }
    // packed lane-link / cross tables
    {
        std::vector<LLInfo> li(K > 0 ? K : 1);
        for (int k = 0; k < K; ++k) {
            li[k].start_lane = s->ll_start_lane[k]; li[k].end_lane = s->ll_end_lane[k];
            li[k].cross_off = s->ll_cross_off[k]; li[k].cross_end = s->ll_cross_off[k + 1];
            li[k].length = s->drv_length[L + k]; li[k].type = s->ll_type[k];
            if (s->ll_signal[k] < 0 || s->ll_signal[k] >= A || s->ll_roadlink[k] < 0 || s->ll_roadlink[k] >= 32) {
                return fail(TSC_EINVAL, "lane-link %d: bad signal / road-link index", k);
            }
            li[k].sigbit = s->ll_signal[k] | (s->ll_roadlink[k] << 16);
        }
        int nx = s->n_cross_entries;
        std::vector<CrossEntry> ce(nx > 0 ? nx : 1);
        for (int x = 0; x < nx; ++x) {
            int f = s->xr_foe_ll[x];
            if (f < 0 || f >= K) { return fail(TSC_EINVAL, "cross %d: bad lane-link index", x); }
            ce[x].dist = s->xr_dist[x]; ce[x].foe_dist = s->xr_foe_dist[x];
            ce[x].foe_len = s->drv_length[L + f]; ce[x].foe_sl_len = s->drv_length[s->ll_start_lane[f]];
            ce[x].foe_ll = f; ce[x].foe_start_lane = s->ll_start_lane[f]; ce[x].foe_end_lane = s->ll_end_lane[f];
            ce[x].foe_type = s->ll_type[f];
        }
        if ((rc = upload(E, li.data(), li.size(), &S.llinfo))) return rc;
        if ((rc = upload(E, ce.data(), ce.size(), &S.cross))) return rc;

    }
    {
        std::vector<double2> lm(D > 0 ? D : 1);
        for (int k = 0; k < D; ++k) lm[k] = make_double2(s->drv_length[k], s->drv_max_speed[k]);
        if ((rc = upload(E, lm.data(), lm.size(), &S.drv_lm))) return rc;
    }
    {   // packed sibling lists for the head look-ahead
        std::vector<int4> sib(L > 0 ? L : 1);
        for (int l = 0; l < L; ++l) {
            const int e0 = s->lane_ll_off[l], n = s->lane_ll_off[l + 1] - e0;
            int4 v = make_int4(n <= 3 ? n : -1, 0, 0, 0);
            if (n > 0 && n <= 3) v.y = L + s->lane_ll[e0];
            if (n > 1 && n <= 3) v.z = L + s->lane_ll[e0 + 1];
            if (n > 2 && n <= 3) v.w = L + s->lane_ll[e0 + 2];
            sib[l] = v;
        }
        if ((rc = upload(E, sib.data(), sib.size(), &S.lane_sib))) return rc;
    }
    // spawn lanes and creation prefix tables
    E->h_is_spawn.assign(L, 0);
    for (int l = 0; l < L; ++l) {      // a lane that spawns in ANY flow set owns a spare slot in every replica
        bool any = false;
        for (int f = 0; f < S.F; ++f) any = any || s->lane_spawn_off[(size_t) f * (L + 1) + l + 1] > s->lane_spawn_off[(size_t) f * (L + 1) + l];
        if (any) { E->h_spawn_lane.push_back(l); E->h_is_spawn[l] = 1; }
    }
    S.n_spawn_lanes = E->n_spawn_lanes = (int) E->h_spawn_lane.size();
    S.spawn_pure = 1;
    for (int k = 0; k < K; ++k) if (E->h_is_spawn[s->ll_end_lane[k]]) S.spawn_pure = 0;
    if (const char *env = getenv("TSC_B200_SPAWN_OVERLAP")) { if (atoi(env) == 0) S.spawn_pure = 0; }
    if ((rc = upload(E, E->h_spawn_lane.data(), E->h_spawn_lane.size(), &S.spawn_lane))) return rc;
    {
        if (E->n_spawn_lanes > 32767) { return fail(TSC_EINVAL, "more than 32767 spawn lanes"); }
        std::vector<short> idx(L > 0 ? L : 1, (short) -1);
        for (int k = 0; k < E->n_spawn_lanes; ++k) idx[E->h_spawn_lane[k]] = (short) k;
        if ((rc = upload(E, idx.data(), (size_t) L, &S.lane_spawn_idx))) return rc;
    }
    const int H2 = S.horizon + 2;
    std::vector<int> ccnt((size_t) S.F * H2, 0);
    std::vector<long long> cent((size_t) S.F * H2, 0);
    for (int f = 0; f < S.F; ++f) {
        const int *row = s->lane_spawn_off + (size_t) f * (L + 1);
        int *cc = ccnt.data() + (size_t) f * H2;
        long long *ce = cent.data() + (size_t) f * H2;
        for (int at = row[0]; at < row[L]; ++at) {
            const int v = s->lane_spawn_vid[at];
            if (v < 0 || v >= N) return fail(TSC_EINVAL, "lane_spawn_vid[%d] out of range", at);
            const int t = s->veh_tick[v];
            if (t < 0 || t > S.horizon) { return fail(TSC_EINVAL, "veh_tick out of horizon"); }
            cc[t + 1] += 1; ce[t + 1] += t;
        }
        for (int t = 1; t <= S.horizon + 1; ++t) { cc[t] += cc[t - 1]; ce[t] += ce[t - 1]; }
    }
    if ((rc = upload(E, ccnt.data(), ccnt.size(), &S.created_cnt))) return rc;
    if ((rc = upload(E, cent.data(), cent.size(), &S.created_enter))) return rc;
    {   // one 16-byte record per spawn-list entry: what handleWaiting needs to put the vehicle on its first lane
        std::vector<int4> rec(N > 0 ? N : 1, make_int4(-1, INT_MAX, 0, 0));
        for (int at = 0; at < s->lane_spawn_off[(size_t) S.F * (L + 1) - 1]; ++at) {
            const int v = s->lane_spawn_vid[at];
            const int rp0 = s->veh_seq_start[v];
            if (rp0 < 1 || rp0 + 1 >= s->n_route_seq) return fail(TSC_EINVAL, "vehicle %d: route cursor out of range", v);
            rec[at] = make_int4(v, s->veh_tick[v], rp0, s->route_seq[rp0 + 1] & 0xFFFF);
        }
        if ((rc = upload(E, rec.data(), rec.size(), &S.spawn_rec))) return rc;
    }
    E->h_route_seq.assign(s->route_seq, s->route_seq + s->n_route_seq);
    E->h_veh_seq_start.assign(s->veh_seq_start, s->veh_seq_start + N);

    int Vcap = vehicle_capacity > 0 ? vehicle_capacity : 1024;
    // room for the holes finished vehicles leave until the next compaction, on top of the running vehicles asked for
    Vcap = (Vcap + HOLE_MAX + 7) & ~7;
    if (Vcap > 32767) { return fail(TSC_EINVAL, "vehicle_capacity above 32767 (blocker slots are 16-bit signed)"); }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    int async_stage = 2, prefetch_next = 1;      // 2: bulk asynchronous copies (TMA) + mbarrier, 1: cp.async, 0: plain vector copies
    if (const char *env = getenv("TSC_B200_ASYNC_STAGE")) { int v = atoi(env); if (v >= 0 && v <= 2) async_stage = v; }
    if (const char *env = getenv("TSC_B200_PREFETCH")) prefetch_next = atoi(env) != 0;
    // Pick the variant from how many working sets fit an SM's shared memory (the register budget follows from the
    // launch bounds): the first of (256 threads x 4 blocks per SM, 64 registers), (192 x 4, 80), (256 x 3, 80), (384 x 2, 80),
    // (256 x 2, 128), (512 x 1) that fits, else the global-memory workspace.  Measured on B200 with this kernel (Hangzhou,
    // B = 4096, profiles/r02e_variant_sweep.txt): 256 x 4 0.749 ms, 192 x 5 (64 registers) 0.754, 192 x 4 0.769, 160 x 5
    // 0.773, 256 x 3 0.859.  TSC_B200_THREADS / TSC_B200_MIN_BLOCKS / TSC_B200_GMEM override.
    bool one_t = S.T == 1;
    if (const char *env = getenv("TSC_B200_ONE_TEMPLATE")) one_t = one_t && atoi(env) != 0;
    const size_t per_sm_bytes = prop.sharedMemPerMultiprocessor;
    int want_nt = 0, want_minb = 0;
    if (const char *env = getenv("TSC_B200_THREADS")) { int v = atoi(env); if (v == 160 || v == 192 || v == 256 || v == 384 || v == 512 || v == 1024) want_nt = v; }
    if (const char *env = getenv("TSC_B200_MIN_BLOCKS")) { int v = atoi(env); if (v >= 2 && v <= 5) want_minb = v; }
    bool force_gmem = false;
    if (const char *env = getenv("TSC_B200_GMEM")) force_gmem = atoi(env) != 0;
    const int cand[9][2] = {{160, 5}, {192, 5}, {256, 4}, {192, 4}, {256, 3}, {384, 2}, {256, 2}, {512, 1}, {1024, 1}};
    bool chosen = false;
    for (int k = 0; k < 9 && !force_gmem && !chosen; ++k) {
        const int nt = cand[k][0], minb = cand[k][1];
        if (want_nt && nt != want_nt) continue;
        if (want_minb && (nt == 256 || nt == 192) && minb != want_minb) continue;
        // on request only (TSC_B200_THREADS / TSC_B200_MIN_BLOCKS): 32 warps at 64 registers, five 160-thread blocks,
        // five 192-thread blocks at 64 registers
        if (!want_nt && (nt == 1024 || nt == 160)) continue;
        if (!want_minb && nt == 192 && minb == 5) continue;
        if (!one_t && !((nt == 256 && minb == 2) || nt == 512)) continue;      // generic-template builds
        if (nt == 160 && !one_t) continue;
        build_layout(E->Y, S, Vcap, nt / 32);
        if ((size_t) E->Y.smem_bytes > prop.sharedMemPerBlockOptin) continue;
        if ((int) (per_sm_bytes / (size_t) (E->Y.smem_bytes + 1024 + 64)) < minb) continue;      // + per-block reserve and the kernel's static shared memory
        E->nt = nt; E->minb = minb; chosen = true;
    }
    if (!chosen) { E->gmem = true; E->nt = 1024; E->minb = 1; build_layout(E->Y, S, Vcap, 32); }
    E->Y.async_stage = async_stage; E->Y.prefetch_next = prefetch_next;
    E->Y.warp_surgery = 1;
    if (const char *env = getenv("TSC_B200_WARP_SURGERY")) E->Y.warp_surgery = atoi(env) != 0;
    // a fixed-capacity build applies when the columns are laid out for exactly the capacity (TSC_B200_FIXED_CAPACITY=0: never)
    int vc = E->Y.Vlay == E->Y.Vcap ? E->Y.Vcap : 0;
    if (const char *env = getenv("TSC_B200_FIXED_CAPACITY")) { if (atoi(env) == 0) vc = 0; }
    E->kern = kernel_for(E->nt, E->minb, false, one_t, E->gmem, vc);
    E->kern_ctl = kernel_for(E->nt, E->minb, true, one_t, E->gmem, vc);
    E->fixed_capacity = E->kern != kernel_for(E->nt, E->minb, false, one_t, E->gmem, 0);
    E->Y.meta_shared = 0;
    if (E->gmem) {
        // the per-drivable / per-signal arrays, the scratch lists and the header of a replica whose vehicle columns do not fit
        // shared memory usually do (16 x 16 grid: 12 480 drivables, ~ 150 KB): randomly read 2-byte entries and the tick's
        // atomic counters are better off there than in L2 (TSC_B200_GMEM_META_SHARED=0: everything in the workspace)
        const size_t need = (size_t) (E->Y.smem_bytes - E->Y.o_cnt) + sizeof(RepHeader);
        bool want = need + 1024 + 64 <= prop.sharedMemPerBlockOptin;
        if (const char *env = getenv("TSC_B200_GMEM_META_SHARED")) want = want && atoi(env) != 0;
        E->Y.meta_shared = want ? 1 : 0;
    }
    const int dyn_smem = E->gmem ? (E->Y.meta_shared ? E->Y.smem_bytes - E->Y.o_cnt + (int) sizeof(RepHeader) : 0) : E->Y.smem_bytes;
    E->dyn_smem = dyn_smem;
    for (int k = 0; k < 2; ++k) {
        step_kernel_t kern = k ? E->kern_ctl : E->kern;
        if (dyn_smem) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, E->nt, dyn_smem));
        if (per_sm < 1) per_sm = 1;
        int grid = prop.multiProcessorCount * per_sm;
        if (grid > n_replicas) grid = n_replicas;
        (k ? E->grid_ctl : E->grid) = grid;
    }
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, E->kern));
    E->regs = fa.numRegs;

    if (E->gmem) {
        const int g = E->grid > E->grid_ctl ? E->grid : E->grid_ctl;
        CUDA_TRY(cudaMalloc((void **) &E->workspace, (size_t) g * (size_t) ((E->Y.smem_bytes + 255) & ~255)));
    }
    CUDA_TRY(cudaMalloc((void **) &E->images, (size_t) (n_replicas + 1) * E->Y.img_bytes));      // + the tick-0 image (tsc_reset_replicas)
    // tick-0 image: empty network, one spare slot per spawn lane
    E->init_image.assign(E->Y.img_bytes, 0);
    {
        RepHeader *h = (RepHeader *) E->init_image.data();
        h->n_slots = 0;
        int *vid = (int *) (E->init_image.data() + E->Y.o_vid);
        for (int i = 0; i < E->Y.Vcap; ++i) vid[i] = -1;
        short *blk = (short *) (E->init_image.data() + E->Y.o_blk);
        for (int i = 0; i < E->Y.Vcap; ++i) blk[i] = -1;
        memset(E->init_image.data() + img_meta(E->Y, E->Y.o_head), 0xFF, 2 * (size_t) (D + 2));      // empty lists
        memset(E->init_image.data() + img_meta(E->Y, E->Y.o_tail), 0xFF, 2 * (size_t) (D + 2));
        memset(E->init_image.data() + E->Y.o_lead, 0xFF, 2 * (size_t) E->Y.Vcap);
        memset(E->init_image.data() + E->Y.o_foll, 0xFF, 2 * (size_t) E->Y.Vcap);
        // the signal programs start on pytsc phase 0 (TSProgram.set_initial_phase, backends/cityflow/traffic_signal.py:26-32):
        // a caller that steps before its first tsc_init_program / action sees that light phase, not raw phase 0
        u8 *sraw = E->init_image.data() + img_meta(E->Y, E->Y.o_sraw);
        for (int a = 0; a < A; ++a) sraw[a] = (u8) s->sig_phase_raw[(size_t) a * s->max_phases];
    }
    size_t io = (size_t) n_replicas * A;
    CUDA_TRY(cudaMalloc((void **) &E->d_actions, io * sizeof(int)));
    CUDA_TRY(cudaMalloc((void **) &E->d_obs, io * S.obs_dim * sizeof(float)));
    CUDA_TRY(cudaMalloc((void **) &E->d_reward, io * sizeof(float)));
    CUDA_TRY(cudaMalloc((void **) &E->d_rg, (size_t) n_replicas * sizeof(float)));
    CUDA_TRY(cudaMalloc((void **) &E->d_mask, io * S.n_actions));
    CUDA_TRY(cudaMallocHost((void **) &E->h_actions, io * sizeof(int)));
    CUDA_TRY(cudaMallocHost((void **) &E->h_obs, io * S.obs_dim * sizeof(float)));
    CUDA_TRY(cudaMallocHost((void **) &E->h_reward, io * sizeof(float)));
    CUDA_TRY(cudaMallocHost((void **) &E->h_rg, (size_t) n_replicas * sizeof(float)));
    CUDA_TRY(cudaMallocHost((void **) &E->h_mask, io * S.n_actions));
    CUDA_TRY(cudaStreamCreateWithFlags(&E->host_compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&E->host_compute2, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&E->host_copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&E->host_ev_actions, cudaEventDisableTiming));
    if (const char *env = getenv("TSC_B200_HOST_STREAMS")) { int v = atoi(env); if (v == 1 || v == 2) E->host_streams = v; }
    for (auto &ev : E->host_ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (const char *env = getenv("TSC_B200_HOST_CHUNKS")) { int v = atoi(env); if (v >= 1 && v <= MAX_HOST_CHUNKS) E->host_chunks = v; }
    if (const char *env = getenv("TSC_B200_HOST_ZERO_COPY")) E->host_zero_copy = atoi(env) != 0;
    if (const char *env = getenv("TSC_B200_HOST_LEAD")) E->host_lead = atoi(env) > 0 ? atoi(env) : 0;
    E->h_flow_set.assign(n_replicas, 0);
    CUDA_TRY(cudaMalloc((void **) &E->d_flow_set, (size_t) n_replicas * sizeof(int)));
    CUDA_TRY(cudaMemset(E->d_flow_set, 0, (size_t) n_replicas * sizeof(int)));
    if ((rc = tsc_reset(E, nullptr))) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

extern "C" {

void tsc_destroy(tsc_handle E) {
    if (!E) return;
    cudaSetDevice(E->device);
    tsc_host_unregister(E);
    cudaFree(E->d_flow_set);
    for (void *p : E->dev_allocs) cudaFree(p);
    cudaFree(E->d_phase_cycles);
    cudaFree(E->workspace);
    if (E->host_compute) cudaStreamDestroy(E->host_compute);
    if (E->host_compute2) cudaStreamDestroy(E->host_compute2);
    if (E->host_copy) cudaStreamDestroy(E->host_copy);
    if (E->host_ev_actions) cudaEventDestroy(E->host_ev_actions);
    for (auto &ev : E->host_ev) if (ev) cudaEventDestroy(ev);
    cudaFree(E->images); cudaFree(E->d_actions); cudaFree(E->d_obs); cudaFree(E->d_reward); cudaFree(E->d_rg); cudaFree(E->d_mask);
    cudaFreeHost(E->h_actions); cudaFreeHost(E->h_obs); cudaFreeHost(E->h_reward); cudaFreeHost(E->h_rg); cudaFreeHost(E->h_mask);
    delete E;
}

int tsc_get_dims(tsc_handle E, int32_t *n_replicas, int32_t *n_lanes, int32_t *n_signals, int32_t *obs_dim, int32_t *state_dim,
                 int32_t *n_actions, int32_t *n_in_total, int32_t *n_out_total, int32_t *visibility) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (n_replicas) *n_replicas = E->B;
    if (n_lanes) *n_lanes = E->S.L;
    if (n_signals) *n_signals = E->S.A;
    if (obs_dim) *obs_dim = E->S.obs_dim;
    if (state_dim) *state_dim = E->S.state_dim;
    if (n_actions) *n_actions = E->S.n_actions;
    if (n_in_total) *n_in_total = E->S.n_in_total;
    if (n_out_total) *n_out_total = E->S.n_out_total;
    if (visibility) *visibility = E->S.visibility;
    return 0;
}

int tsc_reset(tsc_handle E, void *stream) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t st = (cudaStream_t) stream;
    // the initial image is tiny compared with B of them: upload once, then replicate on the device
    CUDA_TRY(cudaMemcpyAsync(E->images + (size_t) E->B * E->Y.img_bytes, E->init_image.data(), E->Y.img_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(E->images, E->init_image.data(), E->Y.img_bytes, cudaMemcpyHostToDevice, st));
    size_t done = 1;
    while (done < (size_t) E->B) {
        size_t n = done < (size_t) E->B - done ? done : (size_t) E->B - done;
        CUDA_TRY(cudaMemcpyAsync(E->images + done * E->Y.img_bytes, E->images, n * E->Y.img_bytes, cudaMemcpyDeviceToDevice, st));
        done += n;
    }
    // every replica's flow set: one strided copy of the [B] table into the image headers
    if (E->S.F > 1)
        CUDA_TRY(cudaMemcpy2DAsync(E->images + offsetof(RepHeader, flow_set), E->Y.img_bytes, E->d_flow_set, sizeof(int), sizeof(int),
                                   E->B, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int tsc_reset_flows(tsc_handle E, const int32_t *flow_set_per_replica, void *stream) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (flow_set_per_replica) {
        for (int b = 0; b < E->B; ++b)
            if (flow_set_per_replica[b] < 0 || flow_set_per_replica[b] >= E->S.F)
                return fail(TSC_EINVAL, "replica %d: flow set %d out of range (the scenario has %d)", b, flow_set_per_replica[b], E->S.F);
        CUDA_TRY(cudaSetDevice(E->device));
        E->h_flow_set.assign(flow_set_per_replica, flow_set_per_replica + E->B);
        CUDA_TRY(cudaMemcpyAsync(E->d_flow_set, E->h_flow_set.data(), (size_t) E->B * sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t) stream));
    }
    return tsc_reset(E, stream);
}

int tsc_reset_replicas_flows(tsc_handle E, const int32_t *replicas, const int32_t *flow_sets, int32_t n, void *stream) {
    if (!E || n < 0 || (n && !replicas)) return fail(TSC_EINVAL, "bad argument");
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t st = (cudaStream_t) stream;
    for (int k = 0; k < n; ++k) {
        if (replicas[k] < 0 || replicas[k] >= E->B) return fail(TSC_EINVAL, "replica index %d out of range", replicas[k]);
        if (flow_sets && (flow_sets[k] < 0 || flow_sets[k] >= E->S.F)) return fail(TSC_EINVAL, "flow set %d out of range (the scenario has %d)", flow_sets[k], E->S.F);
    }
    // the tick-0 image is kept on the device right behind the B replica images
    for (int k = 0; k < n; ++k) {
        const int b = replicas[k];
        CUDA_TRY(cudaMemcpyAsync(E->images + (size_t) b * E->Y.img_bytes, E->images + (size_t) E->B * E->Y.img_bytes,
                                 E->Y.img_bytes, cudaMemcpyDeviceToDevice, st));
        if (flow_sets) E->h_flow_set[b] = flow_sets[k];
        if (E->S.F > 1) {
            if (flow_sets) CUDA_TRY(cudaMemcpyAsync(E->d_flow_set + b, &E->h_flow_set[b], sizeof(int), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(E->images + (size_t) b * E->Y.img_bytes + offsetof(RepHeader, flow_set), E->d_flow_set + b, sizeof(int),
                                     cudaMemcpyDeviceToDevice, st));
        }
    }
    return 0;
}

int tsc_reset_replicas(tsc_handle E, const int32_t *replicas, int32_t n, void *stream) {
    return tsc_reset_replicas_flows(E, replicas, nullptr, n, stream);
}

// blob = StateHeader + B replica images
struct StateHeader { char magic[8]; int32_t abi, B, img_bytes, Vcap, D, A, N, pad; };
static void fill_state_header(tsc_handle E, StateHeader &h) {
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "TSCB200S", 8);
    h.abi = TSC_ABI_VERSION; h.B = E->B; h.img_bytes = E->Y.img_bytes; h.Vcap = E->Y.Vcap; h.D = E->S.D; h.A = E->S.A; h.N = E->S.N;
}

int64_t tsc_state_bytes(tsc_handle E) { return E ? (int64_t) sizeof(StateHeader) + (int64_t) E->B * E->Y.img_bytes : 0; }

int tsc_save_state(tsc_handle E, void *buf, int64_t buf_bytes, void *stream) {
    if (!E || !buf) return fail(TSC_EINVAL, "null argument");
    if (buf_bytes < tsc_state_bytes(E)) return fail(TSC_EINVAL, "state buffer too small: %lld < %lld", (long long) buf_bytes, (long long) tsc_state_bytes(E));
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t st = (cudaStream_t) stream;
    StateHeader h;
    fill_state_header(E, h);
    CUDA_TRY(cudaMemcpyAsync(buf, &h, sizeof h, cudaMemcpyDefault, st));
    CUDA_TRY(cudaMemcpyAsync((char *) buf + sizeof h, E->images, (size_t) E->B * E->Y.img_bytes, cudaMemcpyDefault, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int tsc_load_state(tsc_handle E, const void *buf, int64_t buf_bytes, void *stream) {
    if (!E || !buf) return fail(TSC_EINVAL, "null argument");
    if (buf_bytes < tsc_state_bytes(E)) return fail(TSC_EINVAL, "state buffer too small");
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t st = (cudaStream_t) stream;
    StateHeader h, want;
    CUDA_TRY(cudaMemcpyAsync(&h, buf, sizeof h, cudaMemcpyDefault, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    fill_state_header(E, want);
    if (memcmp(&h, &want, sizeof h) != 0) return fail(TSC_EINVAL, "state blob was saved by a different scenario / batch size / capacity");
    CUDA_TRY(cudaMemcpyAsync(E->images, (const char *) buf + sizeof h, (size_t) E->B * E->Y.img_bytes, cudaMemcpyDefault, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

static int launch(tsc_handle E, const StepArgs &a, void *stream) {
    CUDA_TRY(cudaSetDevice(E->device));
    const bool ctl = a.apply_actions >= 4 || a.decide_only;
    int grid = ctl ? E->grid_ctl : E->grid;
    if (grid > a.B - a.b0) grid = a.B - a.b0;
    if (grid <= 0) return 0;
    (ctl ? E->kern_ctl : E->kern)<<<grid, E->nt, E->dyn_smem, (cudaStream_t) stream>>>(E->S, E->Y, E->images, a);
    E->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static StepArgs blank_args(tsc_handle E) {
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.b0 = 0; a.B = E->B; a.init_program = -1;
    a.phase_cycles = E->d_phase_cycles;
    a.workspace = E->workspace;
    return a;
}

int tsc_set_phase(tsc_handle E, const int32_t *raw_phase, void *stream) {
    if (!E || !raw_phase) return fail(TSC_EINVAL, "null argument");
    StepArgs a = blank_args(E);
    a.set_raw_phase = 1; a.raw_phase = raw_phase;
    return launch(E, a, stream);
}

int tsc_init_program(tsc_handle E, int32_t phase_index, void *stream) {
    if (!E || phase_index < 0) return fail(TSC_EINVAL, "bad argument");
    for (size_t a = 0; a < E->h_sig_n_phases.size(); ++a)
        if (phase_index >= E->h_sig_n_phases[a]) return fail(TSC_EINVAL, "phase index %d: signal %d has %d phases", phase_index, (int) a, E->h_sig_n_phases[a]);
    StepArgs a = blank_args(E);
    a.init_program = phase_index;
    return launch(E, a, stream);
}

int tsc_step(tsc_handle E, int32_t n_ticks, void *stream) {
    if (!E || n_ticks < 0) return fail(TSC_EINVAL, "bad argument");
    StepArgs a = blank_args(E);
    a.n_ticks = n_ticks;
    return launch(E, a, stream);
}

int tsc_retrieve(tsc_handle E, const tsc_outputs_t *out, void *stream) {
    if (!E || !out) return fail(TSC_EINVAL, "null argument");
    StepArgs a = blank_args(E);
    a.do_retrieve = 1; a.out = *out;
    return launch(E, a, stream);
}

// TSC_CTRL_* -> the kernel's apply_actions code; whether the controller reads the caller's actions
static int controller_mode(int controller) {
    switch (controller) {
        case TSC_CTRL_EXTERNAL: return 1;
        case TSC_CTRL_FIXED_TIME: return 2;
        case TSC_CTRL_PHASE_INDEX: return 3;
        case TSC_CTRL_GREEDY: return 4;
        case TSC_CTRL_MAX_PRESSURE: return 5;
        case TSC_CTRL_SOTL: return 6;
        case TSC_CTRL_RANDOM: return 7;
    }
    return 0;
}
static bool controller_needs_actions(int controller) { return controller == TSC_CTRL_EXTERNAL || controller == TSC_CTRL_PHASE_INDEX; }

int tsc_controller_act(tsc_handle E, int32_t controller, int32_t controller_arg, int32_t *actions_out, int32_t *scores_out,
                       void *stream) {
    if (!E || !actions_out) return fail(TSC_EINVAL, "null argument");
    if (!controller_mode(controller) || controller_needs_actions(controller))
        return fail(TSC_EINVAL, "controller %d is not a rule-based controller", controller);
    StepArgs a = blank_args(E);
    a.apply_actions = controller_mode(controller); a.controller_arg = controller_arg;
    a.decide_only = 1; a.ctl_actions = actions_out; a.ctl_scores = scores_out;
    return launch(E, a, stream);
}

int tsc_env_step(tsc_handle E, const int32_t *actions, int32_t controller, int32_t controller_arg, int32_t n_ticks,
                 const tsc_outputs_t *out, void *stream) {
    if (!E || n_ticks < 0) return fail(TSC_EINVAL, "bad argument");
    if (!controller_mode(controller)) return fail(TSC_EINVAL, "unknown controller %d", controller);
    if (controller_needs_actions(controller) && !actions) return fail(TSC_EINVAL, "actions required by controller %d", controller);
    StepArgs a = blank_args(E);
    a.apply_actions = controller_mode(controller);
    a.controller_arg = controller_arg; a.actions = actions; a.n_ticks = n_ticks;
    if (out) { a.do_retrieve = 1; a.out = *out; }
    return launch(E, a, stream);
}

// true when `p` is page-locked host memory CUDA can DMA to directly
static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int tsc_env_step_host(tsc_handle E, const int32_t *actions_host, int32_t controller, int32_t controller_arg, int32_t n_ticks,
                      float *obs_host, float *reward_host, uint8_t *mask_host, float *reward_global_host) {
    if (!E || n_ticks < 0) return fail(TSC_EINVAL, "bad argument");
    if (!controller_mode(controller)) return fail(TSC_EINVAL, "unknown controller %d", controller);
    CUDA_TRY(cudaSetDevice(E->device));
    const size_t A = (size_t) E->S.A, io = (size_t) E->B * A;
    // Replicas are independent, so the batch is cut into chunks: while chunk k+1 is being stepped on
    // the compute stream, chunk k's observations / rewards / masks travel to the host on the copy
    // stream.  Both streams are ordered after whatever the caller queued on the default stream.
    // (two compute streams only without a per-block global workspace: concurrent launches would share it)
    cudaStream_t sc = E->host_compute, sc2 = (E->host_streams > 1 && !E->workspace) ? E->host_compute2 : E->host_compute, sd = E->host_copy;
    CUDA_TRY(cudaEventRecord(E->host_ev[0], 0));
    CUDA_TRY(cudaStreamWaitEvent(sc, E->host_ev[0], 0));
    // page-locked caller buffers are used in place; pageable ones go through the handle's pinned staging
    if (controller_needs_actions(controller)) {
        if (!actions_host) return fail(TSC_EINVAL, "actions required by controller %d", controller);
        const int32_t *src = actions_host;
        if (!is_pinned(actions_host)) { memcpy(E->h_actions, actions_host, io * sizeof(int)); src = E->h_actions; }
        CUDA_TRY(cudaMemcpyAsync(E->d_actions, src, io * sizeof(int), cudaMemcpyHostToDevice, sc));
    }
    struct Out { void *user, *stage; const void *dev; size_t row; bool direct; } outs[4] = {
        {obs_host, E->h_obs, E->d_obs, A * E->S.obs_dim * sizeof(float), false},
        {reward_host, E->h_reward, E->d_reward, A * sizeof(float), false},
        {mask_host, E->h_mask, E->d_mask, A * E->S.n_actions, false},
        {reward_global_host, E->h_rg, E->d_rg, sizeof(float), false}};
    bool all_pinned = true;
    for (Out &x : outs) if (x.user) { x.direct = is_pinned(x.user); all_pinned = all_pinned && x.direct; }
    StepArgs a = blank_args(E);
    a.apply_actions = controller_mode(controller);
    a.controller_arg = controller_arg; a.actions = E->d_actions; a.n_ticks = n_ticks;
    a.do_retrieve = 1;
    if (obs_host) a.out.obs = E->d_obs;
    if (reward_host) a.out.reward = E->d_reward;
    if (mask_host) a.out.mask = E->d_mask;
    if (reward_global_host) a.out.reward_global = E->d_rg;
    if (E->host_zero_copy && all_pinned) {
        // Page-locked buffers are mapped into the device address space (UVA): the kernel's coalesced
        // row stores go straight over PCIe while other replicas are still being stepped -- one launch,
        // no separate copy phase.
        void *dp = nullptr;
        if (obs_host) { CUDA_TRY(cudaHostGetDevicePointer(&dp, obs_host, 0)); a.out.obs = (float *) dp; }
        if (reward_host) { CUDA_TRY(cudaHostGetDevicePointer(&dp, reward_host, 0)); a.out.reward = (float *) dp; }
        if (mask_host) { CUDA_TRY(cudaHostGetDevicePointer(&dp, mask_host, 0)); a.out.mask = (uint8_t *) dp; }
        if (reward_global_host) { CUDA_TRY(cudaHostGetDevicePointer(&dp, reward_global_host, 0)); a.out.reward_global = (float *) dp; }
        int rc = launch(E, a, sc);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(sc));
        return 0;
    }
    // chunk = a whole number of waves of the persistent grid, so that no launch ends on a partly
    // filled wave; at most MAX_HOST_CHUNKS chunks
    // The copies are the longer leg (57 MB per step on the bench workload against < 1 ms of kernel); an optional
    // short lead chunk lets the first copy start earlier (off by default: it cost more than it gave).
    int lead = E->host_lead > 1 ? E->host_lead : ((E->host_lead && E->minb > 1) ? E->grid / E->minb : 0);      // > 1: explicit size
    if (lead >= E->B) lead = 0;
    const int total_waves = (E->B - lead + E->grid - 1) / E->grid;
    const int max_chunks = E->host_chunks - (lead ? 1 : 0) > 0 ? E->host_chunks - (lead ? 1 : 0) : 1;
    int wpc = (total_waves + max_chunks - 1) / max_chunks;
    if (wpc < 1) wpc = 1;
    const int chunk = wpc * E->grid;
    if (sc2 != sc) {   // the second compute stream starts after the actions have arrived
        CUDA_TRY(cudaEventRecord(E->host_ev_actions, sc));
        CUDA_TRY(cudaStreamWaitEvent(sc2, E->host_ev_actions, 0));
    }
    int k = 0;
    for (int b0 = 0; b0 < E->B; ++k) {
        const int len = (k == 0 && lead) ? lead : chunk;
        a.b0 = b0;
        a.B = b0 + len < E->B ? b0 + len : E->B;
        b0 = a.B;
        cudaStream_t st = (k & 1) ? sc2 : sc;
        int rc = launch(E, a, st);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(E->host_ev[1 + k], st));
        CUDA_TRY(cudaStreamWaitEvent(sd, E->host_ev[1 + k], 0));
        const bool last = a.B == E->B;
        for (int o = 0; o < 4; ++o) {
            Out &x = outs[o];
            if (!x.user) continue;
            // the observation rows are the bulk: they leave chunk by chunk; the small outputs leave once
            size_t r0 = o == 0 ? (size_t) a.b0 : 0, nr = o == 0 ? (size_t) (a.B - a.b0) : (size_t) E->B;
            if (o != 0 && !last) continue;
            char *dst = (char *) (x.direct ? x.user : x.stage) + r0 * x.row;
            CUDA_TRY(cudaMemcpyAsync(dst, (const char *) x.dev + r0 * x.row, nr * x.row, cudaMemcpyDeviceToHost, sd));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(sd));
    CUDA_TRY(cudaStreamSynchronize(sc));
    if (sc2 != sc) CUDA_TRY(cudaStreamSynchronize(sc2));
    for (Out &x : outs) if (x.user && !x.direct) memcpy(x.user, x.stage, (size_t) E->B * x.row);
    return 0;
}


int tsc_host_unregister(tsc_handle E) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    HostPath *H = E->hp;
    if (!H) return 0;
    cudaSetDevice(E->device);
    cudaStreamSynchronize(E->host_compute);
    {
        std::lock_guard<std::mutex> lk(H->mu);
        H->quit = true;
    }
    H->cv.notify_all();
    for (auto &t : H->threads) t.join();
    for (int k = 0; k < 2; ++k) if (H->pk[k]) cudaFreeHost(H->pk[k]);
    if (H->flags) cudaFreeHost(H->flags);
    delete H;
    E->hp = nullptr;
    return 0;
}

int64_t tsc_host_packet_bytes(tsc_handle E) { return E ? (int64_t) E->B * (E->S.pk_bytes + 4) : 0; }

int tsc_host_register(tsc_handle E, float *obs_host, float *reward_host, uint8_t *mask_host, float *reward_global_host) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (obs_host && E->S.obs_type != TSC_OBS_LANE_FEATURES)
        return fail(TSC_EINVAL, "the registered host path builds lane_features observation rows only (use tsc_env_step_host)");
    CUDA_TRY(cudaSetDevice(E->device));
    tsc_host_unregister(E);
    HostPath *H = new HostPath();
    E->hp = H;
    H->E = E; H->obs = obs_host; H->reward = reward_host; H->mask = mask_host; H->rg = reward_global_host;
    const DevScn &S = E->S;
    const size_t pk_total = (size_t) E->B * S.pk_bytes;
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(cudaHostAlloc((void **) &H->pk[k], pk_total, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(H->pk[k], 0, pk_total);
        CUDA_TRY(cudaHostGetDevicePointer((void **) &H->pk_dev[k], H->pk[k], 0));
    }
    CUDA_TRY(cudaHostAlloc((void **) &H->flags, (size_t) E->B * sizeof(u32), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(H->flags, 0, (size_t) E->B * sizeof(u32));
    CUDA_TRY(cudaHostGetDevicePointer((void **) &H->flags_dev, H->flags, 0));
    for (int x = 0; x < 256; ++x) {
        uint64_t v = 0;
        for (int k = 0; k < 8; ++k) v |= (uint64_t) ((x >> k) & 1) << (8 * k);
        H->mask_lut[x] = v;
    }
    // the rows as an all-zero packet describes them: static and padding columns from the recipe, dynamic lane
    // values 0, every program on phase 0 -- written once; afterwards only changes are written
    if (obs_host) {
        const int A = S.A, row = S.state_dim;
        std::vector<float> block((size_t) A * row);
        for (int sg = 0; sg < A; ++sg)
            for (int k = 0; k < row; ++k) {
                const size_t at = (size_t) sg * row + k;
                const u32 code = E->h_obs_code[at];
                float v = E->h_obs_static[at];
                if (code) v = ((code & 7) == 4 && ((code >> 4) & 0xFF) == 0) ? 1.0f : 0.0f;
                block[at] = v;
            }
        for (int b = 0; b < E->B; ++b) memcpy(obs_host + (size_t) b * A * row, block.data(), block.size() * sizeof(float));
    }
    H->nthreads = E->host_threads_req > 0 ? E->host_threads_req : host_thread_count();
    const int ngroups = (E->B + HOST_GROUP - 1) / HOST_GROUP;
    if (H->nthreads > ngroups) H->nthreads = ngroups;
    // the workers finish the rows; the caller launches, and lends a hand while it waits in _wait (its short naps there want
    // the same fine timer slack as the workers')
    prctl(PR_SET_TIMERSLACK, 2000UL, 0, 0, 0);
    for (int w = 0; w < H->nthreads; ++w) H->threads.emplace_back(host_worker_main, H);
    return 0;
}

int tsc_env_step_registered_begin(tsc_handle E, const int32_t *actions_host, int32_t controller, int32_t controller_arg, int32_t n_ticks) {
    if (!E || n_ticks < 0) return fail(TSC_EINVAL, "bad argument");
    HostPath *H = E->hp;
    if (!H) return fail(TSC_EINVAL, "tsc_host_register has not been called");
    if (H->pending) return fail(TSC_EINVAL, "a registered step is already in flight (tsc_env_step_registered_wait first)");
    if (!controller_mode(controller)) return fail(TSC_EINVAL, "unknown controller %d", controller);
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t sc = E->host_compute;
    CUDA_TRY(cudaEventRecord(E->host_ev[0], 0));         // ordered after whatever the caller queued on the default stream
    CUDA_TRY(cudaStreamWaitEvent(sc, E->host_ev[0], 0));
    const size_t io = (size_t) E->B * E->S.A;
    if (controller_needs_actions(controller)) {
        if (!actions_host) return fail(TSC_EINVAL, "actions required by controller %d", controller);
        // (the copy is queued, not awaited: a page-locked caller buffer must stay untouched until the launch has read it,
        // which _wait guarantees; ordinary memory is snapshotted into the handle's own page-locked buffer first)
        const int32_t *src = actions_host;
        if (!is_pinned(actions_host)) { memcpy(E->h_actions, actions_host, io * sizeof(int)); src = E->h_actions; }
        CUDA_TRY(cudaMemcpyAsync(E->d_actions, src, io * sizeof(int), cudaMemcpyHostToDevice, sc));
    }
    H->seq += 1;
    if (H->seq == 0) { memset(H->flags, 0, (size_t) E->B * sizeof(u32)); H->seq = 1; }
    H->cur ^= 1;
    StepArgs a = blank_args(E);
    a.apply_actions = controller_mode(controller);
    a.controller_arg = controller_arg; a.actions = E->d_actions; a.n_ticks = n_ticks;
    a.do_retrieve = 1;
    a.pk = H->pk_dev[H->cur]; a.pk_flags = H->flags_dev; a.pk_seq = H->seq;
    H->failed.store(0);
    int rc = launch(E, a, sc);
    if (rc) return rc;
    H->groups_done.store((uint64_t) H->seq << 32, std::memory_order_release);
    H->claim.store((uint64_t) H->seq << 32, std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(H->mu);
        H->job_seq = H->seq;
    }
    H->cv.notify_all();
    H->pending = true;
    return 0;
}

int tsc_env_step_registered_wait(tsc_handle E) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    HostPath *H = E->hp;
    if (!H || !H->pending) return fail(TSC_EINVAL, "no registered step in flight");
    CUDA_TRY(cudaSetDevice(E->device));
    cudaStream_t sc = E->host_compute;
    H->pending = false;
    host_work(H, H->seq);      // whatever has not been handed out yet
    const u32 ngroups = (u32) ((E->B + HOST_GROUP - 1) / HOST_GROUP);
    unsigned spins = 0;
    while ((u32) H->groups_done.load(std::memory_order_acquire) < ngroups && !H->failed.load(std::memory_order_relaxed)) {
        if (++spins < 512) __builtin_ia32_pause();
        else { struct timespec ts = {0, 5000}; nanosleep(&ts, nullptr); }      // (the workers may share this thread's core)
        if ((spins & 0x3FF) == 0 && !H->failed.load() && cudaStreamQuery(sc) != cudaErrorNotReady) {
            // the launch is over (or failed): flags that are still missing will never come
            bool missing = false;
            for (int b = 0; b < E->B && !missing; ++b) missing = __atomic_load_n(&H->flags[b], __ATOMIC_ACQUIRE) != H->seq;
            if (missing) H->failed.store(1);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(sc));
    if (H->failed.load()) return fail(TSC_ECUDA, "registered host step: replica packets did not arrive (code %d)", H->failed.load());
    return 0;
}

int tsc_env_step_registered(tsc_handle E, const int32_t *actions_host, int32_t controller, int32_t controller_arg, int32_t n_ticks) {
    int rc = tsc_env_step_registered_begin(E, actions_host, controller, controller_arg, n_ticks);
    return rc ? rc : tsc_env_step_registered_wait(E);
}

int tsc_host_threads(tsc_handle E, int32_t n) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (n < 0 || n > 64) return fail(TSC_EINVAL, "host thread count out of range");
    E->host_threads_req = n;
    return n > 0 ? n : host_thread_count();
}

int tsc_snapshot(tsc_handle E, int32_t b, int32_t cap, int32_t *vid, int32_t *drivable, double *distance, double *speed,
                 int32_t *blocker_vid, int32_t *enter_ll_time) {
    if (!E || b < 0 || b >= E->B) return fail(TSC_EINVAL, "bad replica index");
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<unsigned char> img(E->Y.img_bytes);
    CUDA_TRY(cudaMemcpy(img.data(), E->images + (size_t) b * E->Y.img_bytes, E->Y.img_bytes, cudaMemcpyDeviceToHost));
    const Layout &Y = E->Y;
    const RepHeader *h = (const RepHeader *) img.data();
    const int *v = (const int *) (img.data() + Y.o_vid);
    const int *el = (const int *) (img.data() + Y.o_ellt);
    const u16 *cnt = (const u16 *) (img.data() + img_meta(Y, Y.o_cnt)), *head = (const u16 *) (img.data() + img_meta(Y, Y.o_head));
    const u16 *foll = (const u16 *) (img.data() + Y.o_foll);
    const short *bl = (const short *) (img.data() + Y.o_blk);
    const double *ps = (const double *) (img.data() + Y.o_pos), *sp = (const double *) (img.data() + Y.o_spd);
    // drivable-major, every drivable's list front to back
    int n = 0;
    for (int d = 0; d < E->S.D; ++d) {
        int i = cnt[d] > 0 ? (int) head[d] : (int) NONE16;
        for (int k = 0; k < (int) cnt[d] && i != (int) NONE16; ++k, i = foll[i]) {
            if (i >= h->n_slots || v[i] < 0) return fail(TSC_EORDER, "replica %d: the list of drivable %d is corrupt", b, d);
            if (n < cap) {
                if (vid) vid[n] = v[i];
                if (drivable) drivable[n] = d;
                if (distance) distance[n] = ps[i];
                if (speed) speed[n] = sp[i];
                if (blocker_vid) blocker_vid[n] = (bl[i] >= 0 && bl[i] < h->n_slots) ? v[bl[i]] : -1;
                if (enter_ll_time) enter_ll_time[n] = d >= E->S.L ? el[i] : INT_MAX;
            }
            ++n;
        }
    }
    return n;
}

int tsc_load_snapshot(tsc_handle E, int32_t b, int32_t n, const int32_t *vid, const int32_t *drivable, const double *distance,
                      const double *speed, const int32_t *route_pos) {
    if (!E || b < 0 || b >= E->B || n < 0) return fail(TSC_EINVAL, "bad argument");
    if (n && (!drivable || !distance || !speed)) return fail(TSC_EINVAL, "null array");
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const Layout &Y = E->Y;
    const int D = E->S.D, L = E->S.L;
    std::vector<unsigned char> img(Y.img_bytes);
    CUDA_TRY(cudaMemcpy(img.data(), E->images + (size_t) b * Y.img_bytes, Y.img_bytes, cudaMemcpyDeviceToHost));
    RepHeader *h = (RepHeader *) img.data();
    u16 *cnt = (u16 *) (img.data() + img_meta(Y, Y.o_cnt)), *head = (u16 *) (img.data() + img_meta(Y, Y.o_head)), *tail = (u16 *) (img.data() + img_meta(Y, Y.o_tail));
    u16 *lead = (u16 *) (img.data() + Y.o_lead), *foll = (u16 *) (img.data() + Y.o_foll);
    int *v = (int *) (img.data() + Y.o_vid), *rp = (int *) (img.data() + Y.o_rpos), *el = (int *) (img.data() + Y.o_ellt);
    u16 *dr = (u16 *) (img.data() + Y.o_drv);
    short *bl = (short *) (img.data() + Y.o_blk);
    u8 *pj = img.data() + Y.o_pj;
    double *ps = (double *) (img.data() + Y.o_pos), *sp = (double *) (img.data() + Y.o_spd);
    if (n > Y.Vcap - HOLE_MAX) return fail(TSC_EOVERFLOW, "snapshot of %d vehicles exceeds vehicle_capacity", n);
    for (int d = 0; d < D; ++d) { cnt[d] = 0; head[d] = tail[d] = (u16) NONE16; }
    int prev = -1;
    for (int i = 0; i < n; ++i) {
        if (drivable[i] < prev || drivable[i] >= D || drivable[i] < 0) return fail(TSC_EINVAL, "snapshot must be drivable-major");
        prev = drivable[i];
    }
    for (int i = 0; i < Y.Vcap; ++i) { v[i] = -1; bl[i] = -1; lead[i] = foll[i] = (u16) NONE16; }
    for (int k = 0; k < n; ++k) {      // slot k = k-th vehicle: every drivable's vehicles front to back
        const int d = drivable[k];
        v[k] = vid ? vid[k] : k;
        ps[k] = distance[k]; sp[k] = speed[k];
        rp[k] = route_pos ? route_pos[k] : 0;
        el[k] = d >= L ? 0 : INT_MAX; dr[k] = (u16) d; pj[k] = 0;
        if (cnt[d] == 0) head[d] = (u16) k;
        else { lead[k] = tail[d]; foll[tail[d]] = (u16) k; }
        tail[d] = (u16) k;
        cnt[d] += 1;
    }
    h->n_slots = n; h->n_running = n;
    CUDA_TRY(cudaMemcpy(E->images + (size_t) b * Y.img_bytes, img.data(), Y.img_bytes, cudaMemcpyHostToDevice));
    return 0;
}

int tsc_max_spanning_tree(tsc_handle E, const double *density_map, double *out, void *stream) {
    if (!E || !density_map || !out) return fail(TSC_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(E->device));
    const int A = E->S.A;
    tsc_mst_kernel<<<E->B, 256, (size_t) A * 16, (cudaStream_t) stream>>>(density_map, out, A);
    E->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tsc_check(tsc_handle E, int32_t *first_bad) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<RepHeader> hs(E->B);
    CUDA_TRY(cudaMemcpy2D(hs.data(), sizeof(RepHeader), E->images, E->Y.img_bytes, sizeof(RepHeader), E->B, cudaMemcpyDeviceToHost));
    for (int b = 0; b < E->B; ++b) {
        if (hs[b].err) {
            if (first_bad) *first_bad = b;
            if (hs[b].err & (ERR_OVERFLOW | ERR_ENT_OVERFLOW))
                return fail(TSC_EOVERFLOW, "replica %d exceeded vehicle_capacity (flags 0x%x)", b, hs[b].err);
            if (hs[b].err & ERR_BAD_PHASE)
                return fail(TSC_EINVAL, "replica %d was given a light phase its signal does not have (flags 0x%x)", b, hs[b].err);
            return fail(TSC_EORDER, "replica %d: a vehicle left its drivable out of order (flags 0x%x)", b, hs[b].err);
        }
    }
    return 0;
}

int tsc_counters(tsc_handle E, int32_t *tick, int32_t *n_running, int32_t *n_finished, int32_t *n_slots) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<RepHeader> hs(E->B);
    CUDA_TRY(cudaMemcpy2D(hs.data(), sizeof(RepHeader), E->images, E->Y.img_bytes, sizeof(RepHeader), E->B, cudaMemcpyDeviceToHost));
    for (int b = 0; b < E->B; ++b) {
        if (tick) tick[b] = hs[b].tick;
        if (n_running) n_running[b] = hs[b].n_running;
        if (n_finished) n_finished[b] = hs[b].n_finished;
        if (n_slots) n_slots[b] = hs[b].n_slots;
    }
    return 0;
}

int64_t tsc_launch_count(tsc_handle E) { return E ? E->launches : 0; }

int tsc_debug_timing(tsc_handle E, int32_t enable, uint64_t *cycles_out, int32_t n) {
    if (!E) return fail(TSC_EINVAL, "null handle");
#ifndef TSC_PHASE_TIMING
    if (enable) return fail(TSC_EINVAL, "this build has no phase timing (compile with -DTSC_PHASE_TIMING: tools/phase_timing.py does)");
#endif
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (cycles_out && E->d_phase_cycles) {
        unsigned long long tmp[PT_N];
        CUDA_TRY(cudaMemcpy(tmp, E->d_phase_cycles, sizeof tmp, cudaMemcpyDeviceToHost));
        for (int k = 0; k < n; ++k) cycles_out[k] = k < PT_N ? tmp[k] : 0;
    } else if (cycles_out) {
        for (int k = 0; k < n; ++k) cycles_out[k] = 0;
    }
    if (enable && !E->d_phase_cycles) CUDA_TRY(cudaMalloc((void **) &E->d_phase_cycles, PT_N * sizeof(unsigned long long)));
    if (enable) CUDA_TRY(cudaMemset(E->d_phase_cycles, 0, PT_N * sizeof(unsigned long long)));
    if (!enable && E->d_phase_cycles) { cudaFree(E->d_phase_cycles); E->d_phase_cycles = nullptr; }
    return PT_N;
}

int tsc_kernel_info(tsc_handle E, int32_t *smem_bytes, int32_t *threads, int32_t *grid, int32_t *regs) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (smem_bytes) *smem_bytes = E->gmem ? E->Y.smem_bytes : E->dyn_smem;
    if (threads) *threads = E->nt;
    if (grid) *grid = E->grid;
    if (regs) *regs = E->regs;
    return 0;
}

int tsc_kernel_variant(tsc_handle E, int32_t *staged, int32_t *global_workspace, int32_t *blocks_per_sm) {
    if (!E) return fail(TSC_EINVAL, "null handle");
    if (staged) *staged = E->fixed_capacity ? 1 : 0;
    if (global_workspace) *global_workspace = E->gmem ? (E->Y.meta_shared ? 2 : 1) : 0;
    if (blocks_per_sm) *blocks_per_sm = E->minb;
    return 0;
}

}  // extern "C"
