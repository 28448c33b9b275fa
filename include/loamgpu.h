/* loamgpu.h — C ABI of the B200-native LiDAR registration hot path for gpsCalibration.
 *
 * The reference has no plugin/FFI API: its seams are the bodies of three ROS nodes
 *   scanRegistration  src/gpsCalibration/src/lidar_slam/loam/scanRegistration.cpp  (SR)
 *   laserOdometry     src/gpsCalibration/src/lidar_slam/loam/laserOdometry.cpp     (LO)
 *   laserMapping      src/gpsCalibration/src/lidar_slam/loam/laserMapping.cpp      (LM)
 * Each entry point below names the reference block (file:line) whose body it replaces.  A maintainer keeps the
 * node's subscribers/publishers and swaps the body for the call.
 *
 * Conventions: plain C, no C++ types, no exceptions across the boundary.  Every function returns 0 (LOAM_OK) or a
 * negative error code.  The caller owns every host pointer it passes; the library owns all device memory and pinned
 * staging inside the handle.  A handle is bound to one CUDA device and one stream and is NOT thread-safe (neither
 * are the reference's nodes); different handles are independent.  Clouds are arrays of float4 {x, y, z, intensity}
 * (16 B/point; the reference's pcl::PointXYZI payload, CH:49); the raw sweep is packed float xyz with a byte stride.
 * Poses are float[6] = {rx, ry, rz, tx, ty, tz} exactly like `transformation[6]` (LO:111) / `transformSum[6]`
 * (LO:112) / `transformTobeMapped[6]` (LM:108).
 */
#ifndef LOAMGPU_H
#define LOAMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOAM_OK 0
#define LOAM_EINVAL (-1) /* bad argument */
#define LOAM_ECUDA (-2)  /* CUDA runtime error (loam_last_cuda_error has the text) */
#define LOAM_ENOSPC (-3) /* caller buffer or internal capacity too small */
#define LOAM_ESTATE (-4) /* call order violated (e.g. odometry before any extract) */
#define LOAM_EUNSUPPORTED (-5) /* reserved (sweeps with empty rings are handled like the reference does, SR:480-490) */

typedef struct loam_handle loam_handle;

/* Compile-time constants of the reference that are structural rather than numeric.  All
 * thresholds (0.1, 0.05, 25, 1.0, 3x, 0.2 ...) stay compile-time constants as in the reference. */
typedef struct loam_params {
  int n_scans;         /* SR:65 N_SCANS = 16 (max 64) */
  int ring_mode;       /* 0: the reference's 16-entry ring table SR:301-320; 1: uniform table (HDL-64-shaped) */
  float ring_ang_min;  /* ring_mode 1: ring = int((angle - ring_ang_min) / ring_ang_step + 0.5) */
  float ring_ang_step;
  int skip_frame_num;  /* LO:52 skipFrameNum = 1: clouds go to mapping every (skip+1)-th sweep */
  int max_points;      /* capacity hint: points per sweep (CH:15 POINTSNUM fence lifted); grows on demand */
  int max_map_points;  /* capacity hint: points of the gathered local map; grows on demand */
  int want_registered; /* 1: mapping also produces /velodyne_cloud_registered (LM:1103-1112) */
  int want_surround;   /* 1: mapping also produces /laser_cloud_surround every mapFrameNum runs (LM:1081-1101) */
  int pose_message_hop; /* 1: loam_process_sweep / the pipeline hand the odometry pose to mapping through loam_pose_message_hop,
                        * like the reference's nodes do through /laser_odom_to_init; 0 (default): the float[6] as it is */
  int gn_max_ctas;     /* 0 (default): the mapping Gauss-Newton loop of a sweep-sized stack uses one CTA per SM; > 0: at most this
                        * many CTAs, so that several sequences sharing one GPU run their loops side by side instead of queueing
                        * for all the SMs (8 pipelines on one B200: 7.4 k -> 8.4 k sweeps/s at 20).  Results do not depend on it. */
} loam_params;

/* cloud selectors for loam_get_cloud */
enum {
  LOAM_CLOUD_FULL = 0,         /* /velodyne_cloud_2          SR:689-693 */
  LOAM_CLOUD_SHARP = 1,        /* /laser_cloud_sharp         SR:698-702 */
  LOAM_CLOUD_LESS_SHARP = 2,   /* /laser_cloud_less_sharp    SR:706-710 */
  LOAM_CLOUD_FLAT = 3,         /* /laser_cloud_flat          SR:714-718 */
  LOAM_CLOUD_LESS_FLAT = 4,    /* /laser_cloud_less_flat     SR:722-726 */
  LOAM_CLOUD_CORNER_LAST = 5,  /* /laser_cloud_corner_last   LO:541-545,1129-1133 (current kd-tree target) */
  LOAM_CLOUD_SURF_LAST = 6,    /* /laser_cloud_surf_last     LO:547-551,1135-1139 */
  LOAM_CLOUD_FULL_RES3 = 7,    /* /velodyne_cloud_3          LO:1141-1145 (valid after a publishing sweep) */
  LOAM_CLOUD_CORNER_STACK = 8, /* laserCloudCornerStack      LM:736-739 */
  LOAM_CLOUD_SURF_STACK = 9,   /* laserCloudSurfStack        LM:741-744 */
  LOAM_CLOUD_CORNER_MAP = 10,  /* laserCloudCornerFromMap    LM:717-722 */
  LOAM_CLOUD_SURF_MAP = 11,    /* laserCloudSurfFromMap      LM:717-722 */
  LOAM_CLOUD_SURROUND = 12,    /* /laser_cloud_surround      LM:1085-1100 */
  LOAM_CLOUD_REGISTERED = 13   /* /velodyne_cloud_registered LM:1103-1112 */
};

/* diagnostic selectors for loam_get_diag (parity tests) */
enum {
  LOAM_DIAG_CURVATURE = 0,   /* float[n_full]  cloudCurvature       SR:475 */
  LOAM_DIAG_PICKED_MASK = 1, /* uint8[n_full]  cloudNeighborPicked after SR:492-549, before selection */
  LOAM_DIAG_LABEL = 2,       /* int8[n_full]   cloudLabel           SR:586-632 */
  LOAM_DIAG_SCAN_START = 3,  /* int32[n_scans] scanStartInd         SR:484,489 */
  LOAM_DIAG_SCAN_END = 4     /* int32[n_scans] scanEndInd           SR:485,490 */
};

typedef struct loam_counts { /* sizes of the five clouds scanRegistration publishes */
  int n_full, n_sharp, n_less_sharp, n_flat, n_less_flat;
} loam_counts;

typedef struct loam_odom_result { /* what one laserOdometry loop body publishes (LO:502-1147) */
  float transform_sum[6];  /* /laser_odom_to_init pose (LO:1059-1064); zeros on the (re)initialisation sweep */
  float transformation[6]; /* sweep-relative transform after the Gauss-Newton loop */
  int odom_published;      /* 0 only on the (re)initialisation sweep (LO:519-563) */
  int clouds_published;    /* corner/surf last published (LO:541-551 or LO:1126-1139) */
  int fullres_published;   /* /velodyne_cloud_3 published (LO:1141-1145): mapping has a full message set */
  int iterations;          /* Gauss-Newton iterations executed (<= 25) */
  int n_corner_last, n_surf_last;
} loam_odom_result;

typedef struct loam_map_result { /* what one laserMapping loop body publishes (LM:425-1139) */
  float transform_aft_mapped[6];  /* /aft_mapped_to_init pose            LM:1114-1124 */
  float transform_bef_mapped[6];  /* carried in the twist fields         LM:1125-1130 */
  float transform_tobe_mapped[6]; /* pose used for the registered cloud  LM:1103-1106 */
  int optimised;                  /* local map large enough (LM:749) */
  int iterations;                 /* Gauss-Newton iterations executed (<= 10) */
  int surround_published;         /* mapFrameCount hit mapFrameNum (LM:1082) */
  int n_corner_stack, n_surf_stack, n_corner_map, n_surf_map;
  int n_surround, n_registered;
} loam_map_result;

typedef struct loam_sweep_result { /* loam_process_sweep: SR -> LO -> LM for one sweep */
  loam_counts counts;
  loam_odom_result odom;
  loam_map_result map; /* valid when mapping_ran */
  int mapping_ran;
} loam_sweep_result;

/* ---- lifecycle -------------------------------------------------------------------------------------------- */
const char* loam_strerror(int code);
const char* loam_last_cuda_error(const loam_handle* h);
void loam_default_params(loam_params* p);
int loam_create(const loam_params* p, int device, loam_handle** out);
int loam_destroy(loam_handle* h);
/* IMControl{systemInited=false} (IN:281-284 -> LO:411-415): the next sweep re-initialises odometry, and mapping
 * resets when it then sees the zero odometry pose (LM:316-319, 434-461). */
int loam_reset(loam_handle* h);
/* cudaStream_t of the handle, for CUDA-event timing by the caller. */
void* loam_stream(loam_handle* h);
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
long long loam_launch_count(const loam_handle* h);
/* out4 = {kernel launches, host->device bytes, device->host bytes, stream synchronisations} since loam_create. */
int loam_stats(const loam_handle* h, long long out4[4]);
/* Optional per-kernel-class CUDA-event timing on the handle's stream (off by default; adds two event records per
 * launch group while on).  Classes: 0 extract, 1 odom_knn, 2 odom_iter, 3 to_end, 4 map_stack/register, 5 voxel,
 * 6 gather, 7 grid build, 8 map_knn, 9 map_fit, 10 insert, 11 sr_select (not counted in 0).  loam_profile(h, 1) clears
 * the counters. */
#define LOAM_PROFILE_CLASSES 12
/* Host wall-clock seconds per section of the node-level calls (incl. waits on the GPU); out16 receives
 * LOAM_HOST_SECTIONS = 16 doubles: 0 extract, 1 odometry iterations, 2 odometry end, 3 mapping prepare (stack transform),
 * 4 grid build, 5 mapping iterations, 6 insert, 7 cube voxel grids, 8 rest, and the finer split of the mapping run:
 * 9 cube bookkeeping / grid roll, 10 gather of the local map, 11 stack voxel grid, 12 arena reserve, 13 cube tables +
 * uploads, 14 cube merge launches, 15 cube merge wait.  Diagnostics only. */
#define LOAM_HOST_SECTIONS 16
int loam_host_times(loam_handle* h, double* out16, int clear);
int loam_profile(loam_handle* h, int enable);
int loam_profile_read(loam_handle* h, double* ms, double* units, long long* scopes, int n);
/* The two latencies every small launch of the library pays on this box, measured on the handle's stream with an empty
 * kernel: period of n back-to-back launches (us per launch) and launch + host-visible completion (us).  bench.py prints
 * the latency floor of one registration (BASELINE configs[0]) from them.  Diagnostics only. */
int loam_launch_latency(loam_handle* h, int n, double* period_us, double* roundtrip_us);

/* ---- scanRegistration: replaces the body of laserCloudHandler, SR:238-752 ------------------------------------
 * xyz: n points in the SENSOR frame (x fwd, y left, z up), `stride_bytes` apart: 12 for packed xyz, 16 for PointXYZ,
 * or the point_step of a sensor_msgs/PointCloud2 payload whose x y z fields are contiguous (pass data + offset of x;
 * 32 for PCL's PointXYZI wire layout, 22 for the Velodyne driver's PointXYZIR: any step >= 12 and any byte alignment is
 * accepted -- this replaces pcl::fromROSMsg at SR:260-261).
 * imu_trans: the 12 floats of /imu_trans (SR:730-745), NULL = zeros (IMU branch dormant in the shipped pipeline).
 * Results stay device-resident for loam_odometry_process; counts are returned; clouds via loam_get_cloud. */
int loam_extract(loam_handle* h, const float* xyz_host, int n, int stride_bytes, double stamp, const float* imu_trans, loam_counts* out);
/* Same, the sweep already resident in device memory of this handle's GPU. */
int loam_extract_device(loam_handle* h, const float* xyz_dev, int n, int stride_bytes, double stamp, const float* imu_trans, loam_counts* out);
/* The IMU branch of scanRegistration (dormant in the shipped pipeline -- input_data replays only velodyne_points -- but
 * live code).  loam_imu_push replaces the body of imuHandler (SR:754-837 incl. AccumulateIMUShift SR:187-233): one
 * sensor_msgs/Imu message (orientation quaternion {x, y, z, w}, angular velocity, linear acceleration, header stamp).
 * From the first message on, loam_extract / loam_extract_device de-skew every kept point of a sweep with the IMU
 * values interpolated at the point's own time (SR:364-434, SR:121-184: a kernel in front of the ring bucketing; `stamp`
 * = the sweep's header stamp, timeScanCur SR:257) and compute /imu_trans themselves (loam_get_imu_trans: the 12 floats
 * as published SR:730-745; they replace a caller-supplied imu_trans and feed the odometry of the same handle).
 * Stated assumption: the stamps in the 200-message ring increase with the ring order (as IMU streams do). */
int loam_imu_push(loam_handle* h, double stamp, const double orientation_xyzw[4], const double angular_velocity[3],
                  const double linear_acceleration[3]);
int loam_get_imu_trans(loam_handle* h, float out12[12]);
/* Batched form of loam_extract (SURVEY 8b `*_batch`): one sweep of each of B independent sequences (B handles on the same
 * device, same n_scans).  Every extraction kernel is launched ONCE for the whole batch (grid.y = sequence) and the counts
 * come back with one synchronisation, so B sweeps cost eight launches instead of 8 B; results are those of B loam_extract
 * calls bit for bit (empty sweeps, point_step not a multiple of 4 and sweeps with empty rings fall back to the per-handle
 * path inside the call).  xyz_host[b] / n[b]: member b's sweep; stamps may be NULL; out[b] receives its counts.  Only the
 * extraction is batched here; loam_odometry_process_batch is the lock-step form of the odometry; the mapping stays per sequence
 * (DESIGN.md §6).  If the call fails part-way the members are in an unspecified state: reset those sequences. */
int loam_extract_batch(loam_handle* const* hs, int B, const float* const* xyz_host, const int* n, int stride_bytes,
                       const double* stamps, loam_counts* out);

/* ---- laserOdometry: replaces the loop body LO:502-1147 for the message set of the last loam_extract ---------- */
int loam_odometry_process(loam_handle* h, loam_odom_result* out);
/* Batched (lock-step) form (SURVEY 8b `*_batch`): the current sweep of B independent sequences (handles on one device, each
 * after its loam_extract / loam_extract_batch).  Every odometry kernel of a round -- box hierarchy, correspondence refresh,
 * iteration 0, a block of device iterations, TransformToEnd -- is launched once for all members taking part (grid.y =
 * sequence); the per-member host steps run between the rounds.  out[b] and the handles' state are those of B
 * loam_odometry_process calls bit for bit. */
int loam_odometry_process_batch(loam_handle* const* hs, int B, loam_odom_result* out);

/* ---- laserMapping: replaces laserOdometryHandler's reset test (LM:316-319) and the loop body LM:425-1139 ------
 * loam_mapping_odometry must be called for every published odometry message (every sweep), loam_mapping_process
 * only when odometry published the full message set (fullres_published). */
/* What the odometry pose goes through between the two reference nodes: laserOdometry publishes transformSum as a
 * quaternion message built from (rz, -rx, -ry) (LO:1066-1078), laserMapping's and transformMaintenance's handlers turn it
 * back with getRPY (LM:322-332, TM:277-283), all in double, including the |pitch| >= pi/2 branch.  An identity up to an
 * ulp; offered so that a caller can be bit-identical to nodes wired through messages.  Host arithmetic. */
int loam_pose_message_hop(const float transform_sum_in[6], float transform_sum_out[6]);
int loam_mapping_odometry(loam_handle* h, const float transform_sum[6]);
int loam_mapping_process(loam_handle* h, loam_map_result* out);

/* ---- transformMaintenance: replaces laserOdometryHandler TM:262-315 and odomAftMappedHandler TM:317-338 --
 * Host-only scalar code (no kernels): fuses every odometry pose with the latest mapping correction and builds the
 * height-compensated planar track.  integrated6 = /integrated_to_init pose; track4 = /true_odometry_to_init
 * {x, y, 10 (HEIGHT), stamp}.  Call loam_integrate_odometry for every published odometry pose and
 * loam_integrate_mapping after every mapping run (transform_aft_mapped, transform_bef_mapped). */
int loam_integrate_odometry(loam_handle* h, const float transform_sum[6], double stamp, float integrated6[6], double track4[4]);
int loam_integrate_mapping(loam_handle* h, const float aft_mapped[6], const float bef_mapped[6]);

/* ---- whole hot path for one sweep (SR -> LO -> LM with device-resident hand-over) -------------- */
int loam_process_sweep(loam_handle* h, const float* xyz_host, int n, int stride_bytes, double stamp, loam_sweep_result* out);
int loam_process_sweep_device(loam_handle* h, const float* xyz_dev, int n, int stride_bytes, double stamp, loam_sweep_result* out);

/* ---- data access ---------------------------------------------------------------------------------------------- */
/* Copies cloud `which` (float4 per point) into host_buf (capacity `cap` points); *n = point count.  LOAM_ENOSPC when
 * cap is too small (*n still set).  host_buf may be NULL to query the size. */
int loam_get_cloud(loam_handle* h, int which, float* host_buf, int cap, int* n);
/* The same cloud as the payload pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZI>) would put on the wire (SR:689-726,
 * LO:1129-1145, LM:1096-1112): point_step 32, x @0, y @4, z @8, 1.0f @12, intensity @16, zero padding @20..31.
 * host_buf holds cap_points * 32 bytes; may be NULL to query the size. */
int loam_get_cloud_wire(loam_handle* h, int which, void* host_buf, int cap_points, int* n_points);
int loam_get_diag(loam_handle* h, int which, void* host_buf, int cap_bytes, int* n_items);

/* ---- stage-level entry points (what the node-level calls are made of; used by the parity tests) --------------- */
/* pcl::VoxelGrid<PointXYZI>::filter (SR:677-683, LM:736-744,1061-1079,1092-1094) on host clouds. */
int loam_voxel_grid(loam_handle* h, const float* in4_host, int m, float leaf, float* out4_host, int cap, int* v);
/* Explicit odometry inputs instead of loam_extract's (tests): current sharp/flat, last corner/surf clouds. */
int loam_odom_set_inputs(loam_handle* h, const float* sharp, int n_sharp, const float* flat, int n_flat,
                         const float* corner_last, int n_corner_last, const float* surf_last, int n_surf_last);
/* One Gauss-Newton iteration body without the solve, LO:586-974: TransformToStart, correspondence refresh when
 * iter % 5 == 0, point-to-line / point-to-plane coefficients, Jacobian rows, AtA (6x6, row-major) and AtB.
 * n_sel < 10 => AtA/AtB zeroed, caller skips the solve (LO:904-907). */
int loam_odom_iter(loam_handle* h, int iter, const float T[6], float AtA[36], float AtB[6], int* n_sel);
/* pointSearchCornerInd1/2, pointSearchSurfInd1/2/3 (LO:102-109) as int32, -1 = none. */
int loam_odom_get_corr(loam_handle* h, int* c1, int* c2, int cap_c, int* s1, int* s2, int* s3, int cap_s);
/* TransformToEnd (LO:156-227) of a host cloud with transform T and imu_trans (NULL = zeros). */
int loam_transform_to_end(loam_handle* h, const float* in4_host, int n, const float T[6], const float* imu_trans, float* out4_host);
/* Explicit mapping inputs (tests): down-sampled stacks (sensor frame) and the gathered local map (map frame). */
int loam_map_set_inputs(loam_handle* h, const float* corner_stack, int n_cs, const float* surf_stack, int n_ss,
                        const float* corner_map, int n_cm, const float* surf_map, int n_sm);
/* One Gauss-Newton iteration body without the solve, LM:754-967: kNN-5, line / plane fit, rows, AtA, AtB.
 * n_sel < 50 => AtA/AtB zeroed (LM:929-932). */
int loam_map_iter(loam_handle* h, int iter, const float T[6], float AtA[36], float AtB[6], int* n_sel);
/* pointSearchInd (LM:760,867) of the last stage-level iteration (loam_map_iter / _partial / _allreduce): 5 int32 per stack
 * point, -1 x5 when the 5th neighbour is not within 1 m.  LOAM_ESTATE after loam_map_optimize (the device loop keeps none). */
int loam_map_get_corr(loam_handle* h, int* corner5, int cap_c, int* surf5, int cap_s);
/* The 6x6 solve + degeneracy handling the host keeps (LO:975-1004 / LM:968-997).  state37 = matP (36) + isDegenerate. */
int loam_gn_solve(const float AtA[36], const float AtB[6], int iter, float eig_threshold, float state37[37], float X[6]);
/* Mapping iterations with the local map sharded over several GPUs: like loam_map_iter but leaves
 * the 28 partial sums {21 upper-triangle AtA, 6 AtB, n_sel} as doubles in device memory at `partial_dev` for an
 * all-reduce by the caller; loam_map_finish_reduced turns the reduced sums into AtA/AtB. */
int loam_map_iter_partial(loam_handle* h, int iter, const float T[6], double* partial_dev28);
int loam_map_finish_reduced(const double reduced28_host[28], float AtA[36], float AtB[6], int* n_sel);
/* The same exchange fused into the reduction kernel (one process per GPU, NVLink peer memory through CUDA IPC): the last
 * CTA of every rank stores its 28 sums into every peer's exchange buffer, raises a flag, waits for the peers' flags and
 * adds the partials in rank order, so every rank's host mailbox receives bit-identical global sums one NVLink round
 * trip after the slowest rank's reduction -- no NCCL launch, no separate read-back.
 *   loam_shard_export   allocate this rank's exchange buffer, return its 64-byte cudaIpcMemHandle_t
 *   loam_shard_connect  handles = world x 64 bytes gathered from all ranks (any transport), in rank order
 *   loam_map_iter_allreduce  loam_map_iter over the union of all ranks' queries; every rank must call it (<= 8 ranks) */
/* The whole Gauss-Newton loop LM:753-1017 for the inputs of loam_map_set_inputs in ONE launch: kNN-5, fit, normal
 * equations, 6x6 solve, pose update and convergence test all run on the device (the eigen-decomposition of iteration 0,
 * LM:970-997, is verified by the host afterwards; a degenerate problem is replayed iteration by iteration).  T is updated
 * in place; *iterations = iterations executed (<= max_iters; the reference uses 10).  On a handle connected with
 * loam_shard_connect every rank calls it with the same T: the ranks exchange their 28 sums inside the kernel each
 * iteration and end with bit-identical poses.
 *   loam_shard_set_slab  owner rule of a sharded map: this rank evaluates the stack points whose MAP-FRAME x (under the
 *                        current pose, re-evaluated every iteration on the device, LM:244-262) lies in [x_lo, x_hi); the
 *                        caller gives every rank the whole stack and the map points of its slab + 1 m halo */
int loam_map_optimize(loam_handle* h, float T[6], int max_iters, int* iterations);
int loam_shard_set_slab(loam_handle* h, float x_lo, float x_hi);
int loam_shard_export(loam_handle* h, unsigned char handle64[64]);
/* Emulation hook for boxes with fewer GPUs than ranks: deposits rank `from_rank`'s 28 partial sums for the NEXT iteration in
 * this handle's exchange buffer and raises its flag, exactly what that peer's kernel would do over NVLink.  Kernels of
 * several ranks must never wait for one another on ONE GPU (nothing guarantees they run at the same time), so the tests
 * run the ranks one after the other and inject the peers' sums. */
int loam_shard_inject(loam_handle* h, int from_rank, const double sums28_host[28]);
int loam_shard_connect(loam_handle* h, const unsigned char* handles, int world, int rank);
int loam_map_iter_allreduce(loam_handle* h, int iter, const float T[6], float AtA[36], float AtB[6], int* n_sel);

/* ---- pipelined mode: the reference's three-process layout (SR | LO | LM connected by queues) on one GPU ----------
 * Three stage threads inside the library, one handle (state + stream) per stage, device-resident hand-over.  Results
 * are identical to loam_process_sweep; sweep k+2 is extracted while k+1 is registered and k is mapped.
 * submit returns as soon as the host buffer may be reused; wait returns the results in submission order (blocking). */
typedef struct loam_pipeline loam_pipeline;
int loam_pipeline_create(const loam_params* p, int device, loam_pipeline** out);
int loam_pipeline_destroy(loam_pipeline* p);
/* IMControl{false}, ordered with the submitted sweeps.  Also ends the current error epoch: a failed sweep makes the
 * remaining sweeps of ITS epoch return that error from loam_pipeline_wait; sweeps submitted after the reset run again. */
int loam_pipeline_reset(loam_pipeline* p);
/* text of the last CUDA error raised inside the pipeline (stage threads or submit); "" if none */
const char* loam_pipeline_last_error(loam_pipeline* p);
int loam_pipeline_submit(loam_pipeline* p, const float* xyz_host, int n, int stride_bytes, double stamp);
/* device-resident sweep: the buffer must stay valid until the sweep's result has been returned */
int loam_pipeline_submit_device(loam_pipeline* p, const float* xyz_dev, int n, int stride_bytes, double stamp);
/* loam_imu_push for a pipeline (imuHandler SR:754-837): applied by the extraction stage in submission order -- sweeps
 * submitted before the message do not see it, sweeps submitted after it do; `stamp` of the submit calls is timeScanCur
 * (SR:257).  /imu_trans travels with the features to the odometry stage (LO:201-225, 566-568, 1053-1064).  Results equal
 * loam_imu_push + loam_process_sweep on one handle. */
int loam_pipeline_imu_push(loam_pipeline* p, double stamp, const double orientation_xyzw[4], const double angular_velocity[3],
                           const double linear_acceleration[3]);
/* One sweep for each of B pipelines (B independent sequences on one device) with the extraction batched: the B sweeps go
 * through loam_extract_batch in the caller's thread (one launch per extraction kernel for the whole batch), every
 * pipeline's own stages do the rest.  Per-pipeline results are those of loam_pipeline_submit; collect them with
 * loam_pipeline_wait on each pipeline.  Do not mix with loam_pipeline_submit calls in flight on the same pipelines. */
int loam_pipeline_submit_batch(loam_pipeline* const* ps, int B, const float* const* xyz_host, const int* n, int stride_bytes,
                               const double* stamps);
/* The same with the scan-to-scan odometry batched as well (loam_extract_batch + loam_odometry_process_batch in the caller's
 * thread, lock-step over the B sequences); mapping and output stages stay per pipeline.  Use it INSTEAD of
 * loam_pipeline_submit on these pipelines (loam_pipeline_reset is fine: the call waits until the reset has arrived). */
int loam_pipeline_submit_lockstep(loam_pipeline* const* ps, int B, const float* const* xyz_host, const int* n, int stride_bytes,
                                  const double* stamps);
int loam_pipeline_wait(loam_pipeline* p, loam_sweep_result* out);
int loam_pipeline_pending(loam_pipeline* p);
/* cudaStream_t of stage `which` (0 extract, 1 odometry, 2 mapping), for CUDA-event timing by the caller */
void* loam_pipeline_stream(loam_pipeline* p, int which);
int loam_pipeline_stats(loam_pipeline* p, long long out4[4]);
/* diagnostic: seconds each stage thread (0 extract, 1 odometry, 2 mapping) spent working on sweeps, i.e. not waiting for its
 * queue or a free hand-over slot; read when the pipeline is idle */
int loam_pipeline_stage_times(loam_pipeline* p, double* out3, int clear);
/* the handle behind stage `which` (0 extract, 1 odometry, 2 mapping, 3 output: holds LOAM_CLOUD_SURROUND when the pipeline was
 * created with want_surround) for loam_host_times / loam_stats / loam_get_cloud while the pipeline is idle; owned by the pipeline */
loam_handle* loam_pipeline_handle(loam_pipeline* p, int which);

/* ---- N1: segment scheduler, replaces the replay loop of input_data.cpp (IN:244-446) -----------------
 * Decides which messages of a bag list are (re)published to the SLAM pipeline, when the pipeline is reset
 * (IMControl{systemInited=false}, IN:281-285, 348-353) and which tracks go out on /slam_track (IN:355-364, 428-441).
 * Pass 0 cuts tracks of `long_distance` metres without overlap, pass 1 tracks of `short_distance` overlapping by
 * `overlap_distance` (IN:257-262; requires long > short > overlap > 0); distances are measured on the odometry the
 * pipeline returns (IN:78-116).  Host-only.  [first_pass, last_pass] = [0, 1] is the reference's run.  A caller may
 * instead run [0, 0] and [1, 1] on two pipelines / GPUs at once; pass 1 started alone lacks the reference's artefact
 * of a one-pose track at its start (the last pose of pass 0 leaks into it through preOdometry, IN:362). */
typedef struct {
  /* publish message `msg_index` (0-based) of bag `bag_index` to the pipeline; report its stamp and, if the pipeline
   * produced /true_odometry_to_init for it, that pose as {x, y, z, stamp} with *odometry_arrived = 1 */
  void (*publish)(void* user, int bag_index, int msg_index, double* stamp, double odometry_xyzt[4], int* odometry_arrived);
  void (*control)(void* user);                                                         /* reset the pipeline */
  void (*slam_track)(void* user, int track_flag, const double* xyzt, int n_points);    /* one IMTrack message */
  void* user;
} loam_replay_callbacks;
typedef struct {
  long long published, lost, resets, tracks;
} loam_replay_stats;
int loam_replay_segments(const int* messages_per_bag, int n_bags, double long_distance, double short_distance,
                         double overlap_distance, int first_pass, int last_pass, const loam_replay_callbacks* cb,
                         loam_replay_stats* stats);

/* ---- N4: track calibration -- weighted rigid alignment of a SLAM track to its GPS (ENU) track -------
 * Replaces the bodies of trackCalibration (gps_calibration/track_calibration.cc, TC), WeightCoeCal
 * (gps_calibration/weight_calculation.cc, WC) and the re-weighting loop of longDisTrackPro
 * (long_distance_track_process.cpp, LD:57-83).  Tracks are n rows of double {x, y, z, t} (COORDXYZT, CH:31-37), fp64 like
 * the reference, same operation order (sequential sums).  Host arithmetic (north_star keeps the trajectory alignment on
 * the host) except loam_track_smooth mode 0: the O(N^2) loop TC:648-674 as one kernel on `device`, every point's sum
 * in the serial loop's order (bit-identical to it); mode 1 is the O(N) closed form on the host (agrees to ~1e-12
 * relative, no GPU needed).  Eigen's JacobiSVD (TC:506) is restated as a two-sided Jacobi SVD (loam_track_svd3).
 * Quirk fence: WC:18-19 / WC:41-42 read element n of an n-element track for the last point; the index is clamped. */
int loam_track_svd3(const double* h9, double* u9, double* s3, double* v9);               /* H = U diag(S) V^T, row-major 3x3 */
int loam_track_speed_weights(const double* slam_xyzt, int n, double* w);                 /* WC:4-27 */
int loam_track_residual_weights(const double* slam_xyzt, const double* enu_xyzt, const double* cal_xyzt, int n,
                                double* w);                                              /* WC:30-78 */
/* constructor + doICP (TC:4-27, 39-201, 366-545, 583-618): T16 = the 4x4 transform (row-major), rotated_xy = n x 2 */
int loam_track_icp(const double* slam_xyzt, const double* enu_xyzt, const double* w, int n, double* T16, double* rotated_xy);
/* doCalibration (TC:631-689): cal_xyzt = n x 4 calibrated ENU track */
int loam_track_smooth(const double* rotated_xy, const double* enu_xyzt, int n, int mode, int device, double* cal_xyzt);
/* one trackCalibration object end to end, as SD:241-243 uses it; T16 may be NULL */
int loam_track_calibrate(const double* slam_xyzt, const double* enu_xyzt, const double* w, int n, int mode, int device,
                         double* cal_xyzt, double* T16);
/* LD:57-83: speed weights, calibration, then `iterations` (MAXITERATOR = 5) rounds of residual re-weighting with the
 * calibrated track as the new source; w_out = the weights merged into /gps_weight (LD:85), cal_xyzt = the last track */
int loam_track_calibrate_long(const double* slam_xyzt, const double* enu_xyzt, int n, int iterations, int mode, int device,
                              double* w_out, double* cal_xyzt);

#ifdef __cplusplus
}
#endif
#endif /* LOAMGPU_H */
