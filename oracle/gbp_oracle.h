/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the RRT-Connect extend path of
 * LiuShenLan/global_body_planner.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the timed CPU baseline.
 * The product (global_body_planner_b200/, include/gbp_b200.h) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against the
 * unmodified reference compiled into oracle/_ref/libgbp_ref.so (this container), and
 * tests/golden/ (.npz files) holds outputs of that reference build (generator: tests/golden/make_golden.py)
 * for the GPU box, where /root/reference does not exist.  The reference's own test suite pins
 * nothing for this path (test/test_global_body_planner.cpp:9 asserts 1+1==2).
 *
 * All arithmetic is fp64 in the reference's source order; build with -ffp-contract=off.
 * Citations are file:line under /root/reference.
 */
#ifndef GBP_ORACLE_H
#define GBP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (include/global_body_planner/planning_utils.h:21-54) */
#define ORC_H_MAX 0.4
#define ORC_H_MIN 0.075
#define ORC_V_MAX 2.0
#define ORC_V_NOM 0.75
#define ORC_P_MAX 1.0
#define ORC_ANG_ACC_MAX 7.0
#define ORC_ROBOT_L 0.3
#define ORC_ROBOT_W 0.3
#define ORC_ROBOT_H 0.05
#define ORC_M_CONST 13.0
#define ORC_G_CONST 9.81
#define ORC_F_MAX 637.0
#define ORC_MU 1.0
#define ORC_T_F_MIN 0.0
#define ORC_T_F_MAX 0.5
#define ORC_KINEMATICS_RES 0.05
#define ORC_BACKUP_RATIO 0.5
#define ORC_NUM_GEN_STATES 6
#define ORC_GOAL_BOUNDS 0.5
#define ORC_MY_PI 3.14159
#define ORC_FLIGHT 0
#define ORC_STANCE 1
#define ORC_FORWARD 0
#define ORC_REVERSE 1
#define ORC_TRAPPED 0
#define ORC_ADVANCED 1
#define ORC_REACHED 2
#define ORC_RRT_STAR_DELTA 3.0 /* rrt_star_connect.h:59 */

/* flag bits reported next to a verdict */
#define ORC_FLAG_OOG 2u /* a terrain probe the reference semantics reach fell outside the grid (UB there) */

typedef struct {
	int nx, ny;
	const double *x, *y;          /* strictly increasing axes */
	const double *z, *dx, *dy, *dz; /* x-major layers [ix*ny + iy] (fast_terrain_map.h:97-118) */
} orc_terrain;

typedef struct {
	long long substates; /* k: sub-states started (isValidState calls) */
	long long lookups;   /* L: getGroundHeight calls */
	long long nanprobes; /* heightIsNan calls */
	unsigned flags;      /* ORC_FLAG_* */
} orc_counters;

/* ---- Philox4x32-10 stream spec (shared, by specification, with the CUDA sampler) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* 53-bit uniforms number `first`..`first+n-1` of (seed, stream, idx, purpose) */
void orc_uniforms(uint64_t seed, uint64_t stream, uint64_t idx, int purpose, int first, int n, double *u);
double orc_det_log(double x);
void orc_det_sincos(double x, double *s, double *c);

/* ---- terrain (src/fast_terrain_map.cpp) */
double orc_ground_height(const orc_terrain *t, double x, double y, unsigned *flags);
int orc_height_is_nan(const orc_terrain *t, double x, double y, unsigned *flags);
void orc_surface_normal(const orc_terrain *t, double x, double y, double n[3], unsigned *flags);

/* ---- primitives (src/planning_utils.cpp) */
void orc_apply_stance(const double s[8], const double a[10], double t, double out[8]);
void orc_apply_flight(const double s[8], double t, double out[8]);
void orc_apply_stance_reverse(const double s[8], const double a[10], double t, double out[8]);
void orc_rotate_grf(const double n[3], const double f[3], double out[3]);
int orc_is_valid_action(const double a[10]);
int orc_is_valid_state(const orc_terrain *t, const double s[8], int phase, orc_counters *c);
int orc_validate_pair(const orc_terrain *t, const double s[8], const double a[10], int direction, int adaptive,
					  double s_new[8], double *t_new, orc_counters *c);
double orc_pose_distance(const double a[8], const double b[8]);
double orc_state_distance(const double a[8], const double b[8]);
double orc_yaw_distance(const double a[8], const double b[8]);

/* ---- samplers (recipes of planning_utils.cpp:379-515, planner_class.cpp:22-148 on the Philox stream) */
void orc_sample_action(uint64_t seed, uint64_t stream, uint64_t idx, const double normal[3], int dir_flag,
					   double dir_thresh, const double s_from[8], const double s_to[8], double a[10]);
void orc_sample_state(const orc_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx, int dir_flag,
					  double dir_thresh, int speed_dir_flag, const double s_from[8], const double s_to[8],
					  double q[8]);

/* ---- tree queries (src/planner_class.cpp:173-200); verts AoS [nv][8] */
int orc_nearest(const double *verts, long long nv, const double q[8], double *dist, int *unique);
long long orc_near(const double *verts, long long nv, const double q[8], double radius, int *ids, long long cap);

/* ---- connect (src/rrt_connect.cpp:20-91) */
int orc_attempt_connect(const orc_terrain *t, const double s_existing[8], const double s[8], int direction,
						int adaptive, double s_new[8], double a_new[10], orc_counters *c);

/* ---- batch drivers (pthreads over disjoint slices) */
void orc_validate_pairs(const orc_terrain *t, long long n, const double *s, const double *a,
						const unsigned char *dir, int adaptive, unsigned char *verdict, unsigned char *flags,
						double *s_new, double *t_new, long long *counters3, int nthreads);
void orc_sample_actions(uint64_t seed, uint64_t stream, uint64_t idx0, long long n, const double normal[3], double *a);
void orc_sample_states(const orc_terrain *t, uint64_t seed, uint64_t stream, uint64_t idx0, long long n, double *q);
void orc_valid_states(const orc_terrain *t, long long n, const double *s, const unsigned char *phase,
					  unsigned char *verdict, unsigned char *flags);

/* ---- Tier-2 planner: iteration-budgeted RRT-Connect / RRT*-Connect on the Philox stream.
 * Follows rrt.cpp:20-102, rrt_connect.cpp:20-120,230-314, rrt_star_connect.cpp:12-75 with the
 * wall-clock budget replaced by `max_iters`. */
typedef struct {
	int k_candidates;  /* actions sampled per extend: 6 = reference (NUM_GEN_STATES) */
	int best_of_k;     /* 0 = first valid in stream order (reference, rrt.cpp:36-50), 1 = closest valid */
	int max_iters;     /* loop iterations of runRRTConnect (each = forward half + reverse half) */
	int max_vertices;  /* per-tree capacity; planning stops (unsolved) when a tree is full */
	int adaptive;      /* state_action_pair_check_adaptive_step_size_flag */
	int rrt_star;      /* 1 = RRTStarConnectClass::extend (choose parent + rewire, delta = 3.0) */
	int post_process;  /* 1 = run postProcessPath on the stitched path (rrt_connect.cpp:139-227) */
	/* the fork's options (rrt.h:186-199, set by RRTClass::set_*; all off in config/params.yaml:16-27) */
	int state_direction_sampling;       /* randomState(terrain, flag, threshold, speed flag, s_from, s_to), rrt_connect.cpp:246-251, :281-286 */
	int state_direction_speed;          /* state_direction_sampling_speed_direction_flag_ */
	int action_direction_sampling;      /* getRandomAction(surf_norm, direction, flag, threshold, s, s_near), rrt.cpp:34, :49 */
	int cost_add_yaw;                   /* path_cost_ = length * w_l + yaw * w_y (rrt_connect.cpp:270-274, :196, :212) */
	double state_direction_threshold, action_direction_threshold;
	double cost_length_weight, cost_yaw_weight;
} orc_plan_params;

typedef struct {
	int solved;
	int iters;          /* iterations consumed */
	int nv_a, nv_b;     /* vertices in the start / goal tree */
	int path_states;    /* states in the returned path (0 if unsolved) */
	int pad;
	double path_length; /* g(Ta.last)+g(Tb.last) (rrt_connect.cpp:269) or post-processed length */
	double path_yaw;
	double path_duration; /* sum of t_s+t_f over the path's actions (rrt_connect.cpp:463-466) */
	long long pair_checks; /* isValidStateActionPair[Reverse] evaluations ("validated actions") */
	long long nn_queries;
	double path_cost;   /* path_cost_ as the reference leaves it: rrt_connect.cpp:270-274, or postProcessPath's sum (:196, :212) */
	long long reserved;
} orc_plan_stats; /* 80 bytes, same layout as gbp_plan_stats */

/* dump of one tree after a run (AoS, ids in insertion order): what the pin against the reference's own loops compares */
typedef struct {
	int cap, n;
	double *states;   /* [cap][8] */
	double *actions;  /* [cap][10]: the action stored with vertex i (GraphClass::actions[i]) */
	int *parent;      /* [cap], -1 for the root */
	double *g, *y;    /* [cap] */
} orc_tree_dump;

int orc_plan(const orc_terrain *t, const double start[8], const double goal[8], uint64_t seed, uint64_t query,
			 const orc_plan_params *p, orc_plan_stats *st, double *path_states, double *path_actions, int path_cap);
/* orc_plan that also copies the two trees out (ta / tb may be NULL) */
int orc_plan_ex(const orc_terrain *t, const double start[8], const double goal[8], uint64_t seed, uint64_t query,
				const orc_plan_params *p, orc_plan_stats *st, double *path_states, double *path_actions, int path_cap,
				orc_tree_dump *ta, orc_tree_dump *tb);
void orc_plan_batch(const orc_terrain *t, long long nq, const double *starts, const double *goals, uint64_t seed,
					uint64_t query0, const orc_plan_params *p, orc_plan_stats *st, int nthreads);
long long orc_interp_path(int n_actions, const double *states, const double *actions, double dt, long long cap,
						  double *out_s, double *out_t, int *out_phase);
double orc_max_curvature(long long n, const double *states);
int orc_post_process_path(const orc_terrain *t, int ns, double *states, double *actions, int adaptive, double stats3[3]);
/* with the yaw-aware cost of the fork: stats3[2] = sum of dl * w_l + dyaw * w_y when cost_add_yaw (rrt_connect.cpp:196, :212) */
int orc_post_process_path_w(const orc_terrain *t, int ns, double *states, double *actions, int adaptive, int cost_add_yaw,
							double w_length, double w_yaw, double stats3[3]);
/* terrain generators of the publisher node (terrain_map_publisher.cpp:34-231, :253-286); see gbp_oracle.c */
extern const double orc_own_map_default_rects[13][6];
void orc_own_map_axes(int n, double start, double res, double *ax);
void orc_own_map_range(const double *xa, int nx, const double *ya, int ny, const double rect[6], int range[4]);
double orc_own_map_draw(uint64_t seed, int rect_no, uint64_t cell, double mu, double delta);
void orc_own_map(uint64_t seed, int x_size, int y_size, double x_start, double y_start, double res, int n_rect,
				 const double *rects, float *elevation, double geom[3]);
void orc_default_map(float *elevation, double geom[3]);

#ifdef __cplusplus
}
#endif
#endif
