/* gbp_b200.h — C ABI of the B200-native RRT-Connect extend path (libgbp_b200.so).
 *
 * Drop-in boundary for the hot path of LiuShenLan/global_body_planner.  The reference has no FFI of
 * its own (plain C++ classes); each entry point below names the reference interface it
 * replaces (file:line under the reference root).  The C++ classes in include/global_body_planner/
 * (same names and signatures as the reference headers) are thin RAII wrappers over these calls; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   State  = 8 doubles {x,y,z,dx,dy,dz,pitch,dpitch}, Action = 10 doubles
 *            {ax_td,ay_td,az_td,ax_to,ay_to,az_to,t_stance,t_flight,apitch_td,apitch_to}
 *            (planning_utils.h:57-62); arrays of them are dense row-major [n][8] / [n][10].
 *   direction: 0 FORWARD, 1 REVERSE; phase: 0 FLIGHT, 1 STANCE (planning_utils.h:43-49).
 *   extend/connect status: 0 TRAPPED, 1 ADVANCED, 2 REACHED (rrt.h:7-9).
 *   Every function returns 0 (GBP_OK) or a negative GBP_E_* code and never throws; the message of
 *   the last failure on the calling thread is gbp_last_error().
 *   Functions without a suffix take HOST pointers (borrowed for the call; copies happen inside).
 *   Functions ending in _dev take DEVICE pointers plus a cudaStream_t passed as void* (NULL = the
 *   legacy default stream), enqueue work and return without synchronising.
 *   All arithmetic that decides a verdict, an index or a propagated state is fp64 without FMA
 *   contraction, in the reference's operation order (bit-exact by construction except for libm
 *   atan2/sin/cos inside isValidState, see GBP_FLAG_NEAR).  There is no CPU fallback.
 */
#ifndef GBP_B200_H
#define GBP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBP_OK 0
#define GBP_E_INVALID (-1)  /* bad argument */
#define GBP_E_CUDA (-2)     /* CUDA runtime error (no device, launch failure, out of memory) */
#define GBP_E_CAPACITY (-3) /* tree or output buffer full */

#define GBP_FORWARD 0
#define GBP_REVERSE 1
#define GBP_FLIGHT 0
#define GBP_STANCE 1
#define GBP_TRAPPED 0
#define GBP_ADVANCED 1
#define GBP_REACHED 2

/* per-candidate flag bits (optional `flags` outputs) */
#define GBP_FLAG_VALID 1u /* copy of the verdict */
#define GBP_FLAG_OOG 2u   /* a terrain probe reached under the reference's sequential semantics fell outside
                             [x_0,x_last) x [y_0,y_last); the reference has undefined behaviour there, this
                             library uses cell-(0,0)-anchored extrapolation */
#define GBP_FLAG_NEAR 4u  /* guard band: a clearance/reach comparison was decided by a margin below 1e-11 m, or a
                             terrain probe fell within 1e-11 m of a grid line.  isValidState is evaluated with
                             ~1e-12 m accuracy (not bit-identical to glibc's atan2/sin/cos), so these are the
                             only candidates whose verdict is not PROVABLY the reference's */

typedef struct gbp_terrain gbp_terrain; /* device-resident FastTerrainMap */
typedef struct gbp_tree gbp_tree;       /* device-resident GraphClass/PlannerClass store */

const char *gbp_last_error(void);
const char *gbp_version(void);
int gbp_device_count(int *count);
int gbp_set_device(int device);

/* ------------------------------------------------------------------------------------- terrain */
/* FastTerrainMap::loadData (fast_terrain_map.cpp:10-28).  Axes must be strictly increasing
 * (GBP_E_INVALID otherwise); layers are x-major [ix*ny + iy] exactly like z_data_[ix][iy]
 * (fast_terrain_map.h:97-118); dx/dy/dz may be NULL (normal (0,0,1)).  Heights are stored on the
 * device as fp32 when every value converts losslessly (always true for the ROS GridMap ingest,
 * fast_terrain_map.cpp:60-66), else as fp64; results are identical either way. */
int gbp_terrain_create(int nx, int ny, const double *x, const double *y, const double *z, const double *dx,
                       const double *dy, const double *dz, gbp_terrain **out);
/* FastTerrainMap::loadDataFromGridMap (fast_terrain_map.cpp:31-91): float layers in grid_map index
 * order [i*ny + j], index (0,0) = largest x and y, cell position = centre + (0.5(n-1) - i)*res. */
int gbp_terrain_create_gridmap(int nx, int ny, double resolution, double centre_x, double centre_y,
                               const float *elevation, const float *dx, const float *dy, const float *dz,
                               gbp_terrain **out);
/* The reference's on-disk terrain format: <directory>/{x,y,z,dx,dy,dz}data.csv, rows = y, columns = x, '#' comment
 * lines skipped (TerrainMapPublisher::loadCSV / loadMapFromCSV, terrain_map_publisher.cpp:289-370).
 * via_gridmap = 0: the values go to FastTerrainMap::loadData as fp64 (x = first row of xdata, y = first column of
 * ydata) — what the oracle harness does.  via_gridmap = 1: the ROS path — layers and the resolution are rounded to
 * float and indexed like the GridMap the publisher fills (:345-369), then FastTerrainMap::loadDataFromGridMap. */
int gbp_terrain_create_csv(const char *directory, int via_gridmap, gbp_terrain **out);
/* The publisher's procedural terrains.  TerrainMapPublisher::createOwnMap (terrain_map_publisher.cpp:34-96, "create"
 * source): a flat x_size x y_size map (reference: 221 x 161 at 0.05 m from (-0.5, -4.0), :36-38) whose axes are the
 * centimetre-rounded accumulation of :46-60, on which n_rect rectangles {x1, y1, x2, y2, mu, delta} are filled in order
 * (changeOwnMapZDataRectangleRandom :148-176 with findXYIndex :178-231; delta <= 0 or NaN mu = the constant fill of
 * changeOwnMapZDataRectangle :129-146); rects = NULL takes the reference's table (:107-127).  The reference draws the
 * truncated Gaussians from a time(0)-seeded engine; here they come from the Philox stream (TERRAIN cell: purpose 3,
 * idx = iy * x_size + ix, stream = rectangle number, Box-Muller pair b = attempts 2b, 2b+1 of the rejection loop,
 * val = z * delta + mu kept iff mu - delta <= val <= mu + delta, mu after 32 rejections), generated on the device.
 * gbp_own_map_layer returns the float elevation layer in grid_map index order and {resolution, centre x, centre y};
 * gbp_terrain_create_own_map feeds it to the GridMap ingest (no dx / dy / dz layers: normals (0, 0, 1)). */
int gbp_own_map_layer(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double resolution, int n_rect,
                      const double *rects, float *elevation, double *geometry3);
int gbp_terrain_create_own_map(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double resolution,
                               int n_rect, const double *rects, gbp_terrain **out);
/* TerrainMapPublisher::createMap (terrain_map_publisher.cpp:253-286, the default source): 12 x 5 m at 0.2 m centred on
 * (4, 0), a 0.1 m disc of radius 0.5 m around (2, 0), normals (0, 0, 1). */
int gbp_terrain_create_default_map(gbp_terrain **out);
void gbp_terrain_destroy(gbp_terrain *t);
int gbp_terrain_dims(const gbp_terrain *t, int *nx, int *ny, int *cell_bytes);
/* which evaluator serves this terrain: uniform_axes = cell edges are computed, not loaded; mixed_precision = the
 * fp32-around-fp64 evaluator applies (no NaN, uniform axes, pitch >= 1 cm, an interior beyond the border zone, and either
 * fp32 cells or — for fp64 maps with |z| <= 4 m — an fp32-rounded texture copy whose rounding is inside the evaluator's
 * guard band; sub-states it cannot decide are re-evaluated in fp64 on the exact grid) */
int gbp_terrain_flags(const gbp_terrain *t, int *uniform_axes, int *mixed_precision);
/* Environment switches read at terrain creation / launch, for A/B measurements only (results are identical):
 * GBP_NO_MIXED=1 keeps the fp64 evaluator on every terrain, GBP_NO_TEX=1 the 4-load fetch, GBP_NO_L2_WINDOW=1 drops the
 * persisting-L2 access window of the height grid. */
/* texture_gather = 1 when the mixed-precision walk fetches each probe's 2x2 cells with one texture gather from a
 * block-linear copy of the height grid (the default for such terrains; GBP_NO_TEX=1 in the environment at creation
 * time keeps the 4-load form, results are identical) */
int gbp_terrain_fetch_path(const gbp_terrain *t, int *texture_gather);
/* FastTerrainMap::getXData / getYData (fast_terrain_map.cpp:216-223) */
int gbp_terrain_axes(const gbp_terrain *t, double *x, double *y);
/* getGroundHeight (:94-132), heightIsNan (:135-157), getSurfaceNormal (:160-213, not renormalised) */
int gbp_ground_height(const gbp_terrain *t, int64_t n, const double *x, const double *y, double *h, uint8_t *flags);
int gbp_height_is_nan(const gbp_terrain *t, int64_t n, const double *x, const double *y, uint8_t *is_nan);
int gbp_surface_normal(const gbp_terrain *t, int64_t n, const double *x, const double *y, double *normal3);

/* ---------------------------------------------------------------------------------- primitives */
/* applyStance (planning_utils.cpp:237-277), applyFlight (:282-306), applyStanceReverse (:324-370);
 * t has n entries.  kind: 0 stance, 1 flight (actions ignored, may be NULL), 2 stance-reverse. */
int gbp_propagate(int kind, int64_t n, const double *states, const double *actions, const double *t, double *out);
/* rotate_grf (planning_utils.cpp:198-231): the Rodrigues rotation taking +z to normal[i], applied to force[i]; n triples */
int gbp_rotate_grf(int64_t n, const double *normal3, const double *force3, double *out3);
/* calculateCurvature (planning_utils.cpp:884-899): three-point curvature of n (x1,y1,x2,y2,x3,y3) sextuples */
int gbp_curvature(int64_t n, const double *points6, double *curvature);
/* isValidAction (planning_utils.cpp:519-556) */
int gbp_valid_actions(int64_t n, const double *actions, uint8_t *verdict);
/* isValidState (planning_utils.cpp:562-635); phase has n entries */
int gbp_valid_states(const gbp_terrain *t, int64_t n, const double *states, const uint8_t *phase, uint8_t *verdict,
                     uint8_t *flags);
int gbp_valid_states_dev(const gbp_terrain *t, int64_t n, const double *states, const uint8_t *phase, uint8_t *verdict,
                         uint8_t *flags, void *stream);
/* poseDistance / stateDistance / stateYawDistance (planning_utils.cpp:106-127, planning_utils.h:133-145);
 * kind 0 / 1 / 2 */
int gbp_distance(int kind, int64_t n, const double *q1, const double *q2, double *out);

/* isValidStateActionPair / isValidStateActionPairReverse through the 6-argument dispatchers
 * (planning_utils.cpp:645-650, 768-773): fixed step (:713-753, :837-876) or adaptive step
 * (:651-712, :774-836).  direction has n entries.  Outputs the reference leaves unwritten are
 * defined: s_new starts as the input state, t_new as 0.  flags / s_new / t_new may be NULL.
 * `variant` selects the kernel: 0 = default (currently the refill kernel), 1 = one thread per
 * action, 2 = one warp per action / one lane per sub-state (fixed step only), 3 = lane-per-action
 * with warp-level refill (two launches: the walk, then k_pair_outputs which turns the walk's output recipes
 * into exact s_new values).  Variant 5 runs the walk of variant 3 alone and leaves the recipes {tau, kind}
 * in s_new[i][0..1]; gbp_pair_outputs_dev finishes them (bench.py times the two kernels separately this
 * way).  All variants return identical results. */
int gbp_validate_pairs(const gbp_terrain *t, int64_t n, const double *states, const double *actions,
                       const uint8_t *direction, int adaptive, int variant, uint8_t *verdict, uint8_t *flags,
                       double *s_new, double *t_new);
int gbp_validate_pairs_dev(const gbp_terrain *t, int64_t n, const double *states, const double *actions,
                           const uint8_t *direction, int adaptive, int variant, uint8_t *verdict, uint8_t *flags,
                           double *s_new, double *t_new, void *stream);
int gbp_pair_outputs_dev(int64_t n, const double *states, const double *actions, double *s_new, void *stream);
/* per-launch work counters of the last gbp_validate_pairs[_dev] call on this terrain handle, summed
 * over candidates under the REFERENCE's early-exit semantics: {sub-states k, getGroundHeight calls L,
 * heightIsNan calls, candidates flagged OOG, candidates flagged NEAR, valid}.  Synchronises. */
int gbp_validate_counters(const gbp_terrain *t, int64_t counters6[6]);

/* ------------------------------------------------------------------- sample + validate (narrow wire) */
/* The unit of work of RRTClass::newConfig (rrt.cpp:34-50): getRandomAction(surf_norm, ...) followed by
 * isValidStateActionPair[Reverse](s_near, a_test, ...), for n candidates per call, with a wire format that carries
 * only what newConfig exchanges with its caller:
 *   in   the start state of candidate i is a ROW of a device-resident state table (the tree vertices candidates start
 *        from, uploaded once with gbp_states_create): state_idx[i] (4 bytes; NULL = row row0 + i), and direction[i]
 *        (1 byte; NULL = direction0 for all).  Its action is ACTION cell idx0 + i of the Philox stream (seed, stream)
 *        — the same cells gbp_sample_actions(seed, stream, idx0, n, normal, ...) returns — sampled inside the kernel.
 *   out  verdict_bits: bit (i & 31) of word i >> 5 (ceil(n / 32) words; every bit is written) and, for the VALID
 *        candidates only, in ascending candidate order, row j of valid_index / valid_s_new / valid_t_new /
 *        valid_action: the candidate number, the s_new and t_new the pair check returns for it (the landing state and
 *        t_s + t_f, or in REVERSE the exact start state and t_s) and its sampled action — everything rrt.cpp:44-62
 *        reads (s_test of an invalid pair check is never used there).  At most valid_cap rows are written;
 *        result->n_valid is the total.  Any of the valid_* arrays and flags (n bytes, GBP_FLAG_* per candidate) may
 *        be NULL.
 * Verdicts, rows and work counters equal those of gbp_validate_pairs on the same (state, action, direction) triples.
 * Against its 145 B in + 74 B out per candidate this call moves 5 B (4 B with direction_in_row) in + 1 bit out (+ 156 B
 * per valid candidate).
 * All scratch is allocated per call on the call's stream: concurrent calls on one terrain handle are safe. */
typedef struct gbp_states gbp_states; /* device-resident table of start states, rows of 8 doubles */
int gbp_states_create(int64_t rows, const double *states, gbp_states **out);
void gbp_states_destroy(gbp_states *s);
int gbp_states_rows(const gbp_states *s, int64_t *rows);

typedef struct {
	uint64_t seed, stream, idx0;       /* candidate i samples ACTION cell idx0 + i of (seed, stream) */
	double normal[3];                  /* terrain.getSurfaceNormal at the target sample (rrt.cpp:25) */
	int adaptive;                      /* state_action_pair_check_adaptive_step_size_flag_ (rrt.h:186-188) */
	int direction0;                    /* direction of every candidate when `direction` is NULL */
	int action_direction_sampling;     /* getRandomAction(surf_norm, direction, flag, threshold, s, s_near) */
	double action_direction_threshold; /*   (planning_utils.cpp:379-391): s_near = the candidate's start state, */
	double target[8];                  /*   s = target (read only when the flag is set) */
	int64_t row0;                      /* state_idx == NULL: candidate i starts from table row row0 + i */
	int start_states_valid;            /* 1 = the caller promises that every table row is a valid STANCE state, as tree vertices
	                                      are by construction (only end states of fully valid pair checks are ever added,
	                                      rrt.cpp:87, rrt_connect.cpp:110; a root is checked once with gbp_valid_states).  The
	                                      reference re-checks the start state as the first sub-state of EVERY candidate
	                                      (planning_utils.cpp:718-730, :842-848); with the promise that check is counted
	                                      (work counters stay the reference's) but not repeated.  0 = evaluate it. */
	int direction_in_row;              /* 1 = state_idx[i] carries candidate i's direction in bit 31 (0 FORWARD, 1 REVERSE) above a 31-bit
	                                      row number, and `direction` is NULL: ONE 4-byte word per candidate on the wire
	                                      instead of 5 bytes in two arrays.  0 = rows and directions as separate arrays. */
} gbp_sv_params;

typedef struct {
	int64_t n_valid;                             /* valid candidates of the call (may exceed valid_cap) */
	int64_t substates, lookups, nanprobes;       /* work under the reference's early-exit semantics (k, L, heightIsNan calls) */
	int64_t oog, near;                           /* candidates flagged GBP_FLAG_OOG / GBP_FLAG_NEAR */
	int64_t reserved[2];                         /* [0]: row numbers outside the table (results are void when > 0); [1]: internal (0) */
} gbp_sv_result; /* 64 bytes */

/* HOST pointers.  Row numbers and directions travel in blocks on a copy stream while ONE launch walks the call: its warps
 * wait for the block their candidates lie in (after the first block the copies run ahead of the walk); verdict bits come
 * back as one copy, the rows of the valid candidates once their number is known.  GBP_SV_TRACE=1 prints the phases. */
int gbp_sample_validate(const gbp_terrain *t, const gbp_states *table, int64_t n, const int32_t *state_idx,
                        const uint8_t *direction, const gbp_sv_params *params, uint32_t *verdict_bits, uint8_t *flags,
                        int64_t valid_cap, int32_t *valid_index, double *valid_s_new, double *valid_t_new,
                        double *valid_action, gbp_sv_result *result);
/* device pointers throughout (states_dev = the table, [table_rows][8], 16-byte aligned; result_dev = 8 int64 words laid
 * out as gbp_sv_result); enqueues on `stream` and returns.  Row numbers are checked where they are read: one outside
 * [0, table_rows) is read as row 0 and counted in reserved[0] (the host-pointer call returns GBP_E_INVALID then). */
int gbp_sample_validate_dev(const gbp_terrain *t, const double *states_dev, int64_t table_rows, int64_t n, const int32_t *state_idx_dev,
                            const uint8_t *direction_dev, const gbp_sv_params *params, uint32_t *verdict_bits_dev,
                            uint8_t *flags_dev, int64_t valid_cap, int32_t *valid_index_dev, double *valid_s_new_dev,
                            double *valid_t_new_dev, double *valid_action_dev, int64_t *result_dev, void *stream);

/* measurement aid: the walk kernels of the call above alone (no compaction, no rows): verdict bits and 8 counter words
 * {sub-states k, lookups L, NaN probes, OOG, NEAR, valid, rows out of range, 0} — bench.py times the dominant kernel so */
int gbp_sample_validate_walk_dev(const gbp_terrain *t, const double *states_dev, int64_t table_rows, int64_t n,
                                 const int32_t *state_idx_dev, const uint8_t *direction_dev, const gbp_sv_params *params,
                                 uint32_t *verdict_bits_dev, int64_t *counters8_dev, void *stream);

/* ---------------------------------------------------------------------------------- plan output */
/* getInterpPath / interpStateActionPair (planning_utils.cpp:142-193): n_actions primitives, n_actions + 1 states.
 * Per primitive: stance samples for (t = 0; t < t_s; t += dt) (phase 1 STANCE, or 2 CONNECT_STANCE when t_f == 0),
 * flight samples for (t = 0; t < t_f; t += dt) (phase 0 FLIGHT) from the exact take-off state, the exact landing
 * state when t_f > 0 (phase 1), and after the last primitive the final state of the sequence.  *count receives the
 * number of interpolated states (it may exceed cap: only cap entries are written then).  As in the reference the
 * closing state has no phase entry: interp_phase holds *count - 1 values. */
int gbp_interp_path(int n_actions, const double *states, const double *actions, double dt, int64_t cap,
                    double *interp_states, double *interp_t, int *interp_phase, int64_t *count);
/* calculateMaxCurvature (planning_utils.cpp:884-909) over n plan states */
int gbp_max_curvature(int64_t n, const double *states, double *max_curvature);

/* ------------------------------------------------------------------------------------ samplers */
/* Philox4x32-10 stream spec (shared with the CPU harness, oracle/gbp_oracle.c):
 *   key = (seed lo, seed hi); counter = (idx lo, idx hi, stream lo, (stream hi & 0xffffff) | block<<24 | purpose<<28)
 *   uniform j = words (2(j&1), 2(j&1)+1) of block j/2:  u = ((w_hi>>5)*2^26 + (w_lo>>6)) * 2^-53
 * purpose 1 = ACTION cell, 2 = STATE cell.  Candidate i of a call uses idx = idx0 + i. */
/* getRandomAction (planning_utils.cpp:379-442) and getRandomActionDirection (:443-515).
 * normal3: one surface normal for the whole batch (rrt.cpp:25).  s_from/s_to: NULL disables
 * directional sampling; otherwise one State each and dir_threshold as in rrt.h:198-199. */
int gbp_sample_actions(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, const double *normal3,
                       const double *s_from, const double *s_to, double dir_threshold, double *actions);
int gbp_sample_actions_dev(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, const double *normal3_host,
                           double *actions, void *cuda_stream);
/* PlannerClass::randomState (planner_class.cpp:22-76) and randomStateDirection (:82-148) */
int gbp_sample_states(const gbp_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n,
                      const double *s_from, const double *s_to, double dir_threshold, int speed_direction,
                      double *states);
int gbp_sample_states_dev(const gbp_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n,
                          double *states, void *cuda_stream);

/* --------------------------------------------------------------------------------- tree store */
/* GraphClass / PlannerClass storage (graph_class.h:155-170) as a device SoA arena: vertex
 * components in 8 arrays of `capacity` doubles, actions in 10, plus parent, g and yaw-sum.
 * Vertex ids are dense 0..n-1 in insertion order, as rrt.cpp:87 allocates them. */
int gbp_tree_create(int capacity, gbp_tree **out);
void gbp_tree_destroy(gbp_tree *T);
int gbp_tree_init(gbp_tree *T, const double *root_state);            /* GraphClass::init (graph_class.cpp:140-152) */
int gbp_tree_size(const gbp_tree *T, int *n);                         /* getNumVertices (:23-25) */
/* addVertex + addEdge + addAction + updateGYValue as rrt.cpp:87-92; returns the new id */
int gbp_tree_append(gbp_tree *T, int parent, const double *state, const double *action, int *new_id);
/* bulk load of n vertices (ids 0..n-1); parent[0] = -1.  g/yaw are rebuilt as addEdge would. */
int gbp_tree_load(gbp_tree *T, int n, const double *states, const double *actions, const int *parent);
/* read back n vertices starting at id first; any output may be NULL */
int gbp_tree_read(const gbp_tree *T, int first, int n, double *states, double *actions, int *parent, double *g,
                  double *yaw);
/* PlannerClass::getNearestNeighbor (planner_class.cpp:185-200) for m queries.  Ties: lowest id
 * (the reference's order is that of std::unordered_map). */
int gbp_nearest(const gbp_tree *T, int64_t m, const double *queries, int *idx, double *dist);
int gbp_nearest_dev(const gbp_tree *T, int64_t m, const double *queries, int *idx, double *dist, void *stream);
/* PlannerClass::neighborhoodDist (planner_class.cpp:173-182): ids with 0 < d <= radius in ascending
 * id order; *count receives the total (may exceed cap, in which case only cap ids are written). */
int gbp_near(const gbp_tree *T, const double *query, double radius, int *ids, int cap, int *count);

/* -------------------------------------------------------------------------- extend / connect */
/* RRTClass::extend (rrt.cpp:77-102) with newConfig (rrt.cpp:20-70) generalised to K candidates:
 * nearest neighbour, K actions from ACTION cells idx0..idx0+K-1 of (seed, stream), pair checks from
 * s_near in `direction`, selection (best_of_k 0: first valid in stream order decides, the reference's
 * behaviour with K = 6; 1: closest valid), acceptance iff closer to `target` than s_near, append.
 * action_direction_threshold: getRandomAction(surf_norm, direction, flag, threshold, s, s_near) (rrt.cpp:34,
 * planning_utils.cpp:379-391) with s = target; a NEGATIVE value is flag = false (the draw p in [0, 1) is never <= it).
 * One launch; everything stays on the device.  *status, *new_id, checks may be NULL.  A tree that is full when the
 * accepted vertex is to be appended returns GBP_E_CAPACITY (the tree is unchanged), as gbp_connect does. */
int gbp_extend(gbp_tree *T, const gbp_terrain *t, const double *target, int direction, int k_candidates,
               int best_of_k, int adaptive, double action_direction_threshold, uint64_t seed, uint64_t stream,
               uint64_t idx0, int *status, int *new_id, int64_t *pair_checks);
/* RRTClass::newConfig (rrt.cpp:20-70) from an explicit s_near (no tree): K candidates, selection and the
 * "closer than s_near" acceptance as in gbp_extend.  *found = 1 when s_new / a_new were written. */
int gbp_new_config(const gbp_terrain *t, const double *target, const double *s_near, int direction, int k_candidates,
                   int best_of_k, int adaptive, double action_direction_threshold, uint64_t seed, uint64_t stream,
                   uint64_t idx0, int *found, double *s_new, double *a_new, int64_t *pair_checks);
/* RRTConnectClass::attemptConnect (rrt_connect.cpp:20-91), n independent (s_existing, s) pairs.  t_s may be
 * NULL (stance time = poseDistance / V_NOM, the 6-argument overload :85-91) or hold n explicit stance times
 * (the 7-argument overload :20-84). */
int gbp_attempt_connect(const gbp_terrain *t, int64_t n, const double *s_existing, const double *s,
                        const uint8_t *direction, int adaptive, int *status, double *s_new, double *a_new,
                        uint8_t *flags);
int gbp_attempt_connect_ts(const gbp_terrain *t, int64_t n, const double *s_existing, const double *s, const double *t_s,
                           const uint8_t *direction, int adaptive, int *status, double *s_new, double *a_new,
                           uint8_t *flags);
/* RRTConnectClass::connect (rrt_connect.cpp:98-120) */
int gbp_connect(gbp_tree *T, const gbp_terrain *t, const double *target, int direction, int adaptive, int *status,
                int *new_id);

/* ------------------------------------------------------------------------------ batch planner */
/* runRRTConnect (rrt_connect.cpp:230-314) / the RRT*-Connect loop (rrt_star_connect.cpp:130-165)
 * for nq independent queries resident on the device, with the wall-clock budget replaced by an
 * iteration budget.  STATE cell idx = 2*iter + half, ACTION cells idx = (2*iter + half)*K + j,
 * stream = query0 + i. */
typedef struct {
	int k_candidates; /* actions per extend; 6 = NUM_GEN_STATES (planning_utils.h:48) */
	int best_of_k;    /* 0 first valid (reference), 1 closest valid */
	int max_iters;
	int max_vertices; /* per-tree capacity */
	int adaptive;
	int rrt_star;     /* RRTStarConnectClass::extend (rrt_star_connect.cpp:12-75), delta = 3.0 */
	int post_process; /* postProcessPath (rrt_connect.cpp:139-227) on solved queries */
	int stop_after_solved; /* > 0: anytime use (many attempts at ONE query): once this many queries of the batch have
	                          solved, the others stop at their next iteration and report solved = 0 with the work done
	                          so far (which ones depends on timing).  0: every query runs to its own budget (reproducible) */
	/* the fork's options, as RRTClass::set_* stores them (rrt.h:124-139, :186-199; all off in config/params.yaml:16-27) */
	int state_direction_sampling;  /* randomState(terrain, flag, threshold, speed flag, s_from, s_to): rrt_connect.cpp:246-251, :281-286 */
	int state_direction_speed;     /* state_direction_sampling_speed_direction_flag_ */
	int action_direction_sampling; /* getRandomAction(surf_norm, direction, flag, threshold, s, s_near): rrt.cpp:34, :49 */
	int cost_add_yaw;              /* path_cost = length * w_length + yaw * w_yaw (rrt_connect.cpp:270-274, :196, :212) */
	double state_direction_threshold, action_direction_threshold;
	double cost_length_weight, cost_yaw_weight;
} gbp_plan_params; /* 80 bytes */

typedef struct {
	int solved, iters, nv_a, nv_b, path_states, pad;
	double path_length, path_yaw, path_duration;
	int64_t pair_checks, nn_queries;
	double path_cost; /* path_cost_ as the reference leaves it (rrt_connect.cpp:270-274, or postProcessPath's sum :196, :212) */
	int64_t reserved;
} gbp_plan_stats; /* 80 bytes; what the final NCCL gather carries per query */

int gbp_plan_batch(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed,
                   uint64_t query0, const gbp_plan_params *params, gbp_plan_stats *stats, double *path_states,
                   double *path_actions, int path_cap);
int gbp_plan_batch_dev(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed,
                       uint64_t query0, const gbp_plan_params *params, gbp_plan_stats *stats, double *path_states,
                       double *path_actions, int path_cap, void *stream);
/* Which form of the batch planner a call with these parameters and this many queries takes: 0 = the megakernel (k_plan_batch:
 * one warp runs one query's whole search), 1 = the pipelined form (rounds of prep / flattened candidate walk / select /
 * connect kernels over all queries: large batches of plain fixed-step RRT-Connect on terrains with the mixed-precision
 * evaluator), 2 = the stepped form (GBP_PLAN_MODE=step), 3 = the device-wide form (k_plan_wide: one search at a time on a
 * cooperative grid, the candidates of a newConfig spread over all SMs — batches with fewer queries than a newConfig has
 * candidates / 256, e.g. ONE search at K = 4096: 18 us per extend + connect instead of 1.3 ms on one warp; plain fixed-step
 * RRT-Connect).  Results are bit-identical; GBP_PLAN_MODE=mega|pipe|step|wide in the environment forces a form where it
 * applies (A/B measurements). */
int gbp_plan_batch_form(const gbp_terrain *t, const gbp_plan_params *params, int64_t nq, int *form);
/* gbp_plan_batch that also returns every query's two trees as they stand when its search ends (inspection and parity
 * tests: the trees are compared vertex by vertex with the reference's): row ((q * 2 + w) * tree_cap + i) of tree_states
 * [8], tree_actions [10], tree_parent, tree_g, tree_yaw holds vertex i of tree w (0 start side, 1 goal side) of query q;
 * stats[q].nv_a / nv_b give the vertex counts.  HOST pointers. */
int gbp_plan_batch_trees(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed,
                         uint64_t query0, const gbp_plan_params *params, gbp_plan_stats *stats, double *path_states,
                         double *path_actions, int path_cap, int tree_cap, double *tree_states, double *tree_actions,
                         int *tree_parent, double *tree_g, double *tree_yaw);

#ifdef __cplusplus
}
#endif
#endif /* GBP_B200_H */
