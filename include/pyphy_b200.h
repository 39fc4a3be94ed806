/*
 * pyphy_b200.h -- C ABI of the B200-native batched billiards engine.
 *
 * The reference (pulkitag/pyphy-engine) has no FFI: its "plugin surface" for the
 * simulate-and-render path is the Python object API of primitives.py /
 * dynamics.py / dataio.py.  This header is the boundary a binding for that API
 * sits on (ctypes stub in INTEGRATION.md; the in-tree binding is
 * pyphy_engine_b200/_capi.py).  Each entry point names the reference interface
 * it replaces (file:line in the reference tree).
 *
 * Model: one `ppb_batch` holds structure-of-arrays state for `n_worlds`
 * independent worlds in HBM.  All worlds of a batch share one object layout:
 * `n_walls` wall slots (each a closed 4-vertex polygon = 4 collision edges,
 * geometry.py:471-505) and `n_balls` ball slots, plus the insertion order of
 * those objects (primitives.py:936-948), which is the reference's tie-break order
 * (primitives.py:724-733).  A ball slot with radius < 0 is empty.
 *
 * Arithmetic: PPB_F64 follows the reference's fp64 operation order literally
 * (collision events reproduce the reference's bit for bit); PPB_F32 is the
 * throughput mode (same algorithm and quirks, re-associated fp32 math).
 *
 * Threading / streams: a batch is used from one host thread at a time.  Every
 * call enqueues its work on the `stream` argument (a cudaStream_t passed as
 * void*, NULL = the legacy default stream) and, unless its name ends in
 * `_async`, returns after the work has completed.  Host buffers are plain C
 * arrays, world-major and C-contiguous; `*_dev` arguments are device pointers
 * owned by the caller (e.g. torch tensors), never freed by the library.
 *
 * Every function returns PPB_OK (0) or a negative PPB_E_* code;
 * ppb_last_error() gives the message for the calling thread.
 */
#ifndef PYPHY_B200_H
#define PYPHY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPB_VERSION 3

#if defined(__GNUC__)
#define PPB_API __attribute__((visibility("default")))
#else
#define PPB_API
#endif

/* return codes */
#define PPB_OK 0
#define PPB_E_INVALID (-1) /* bad argument */
#define PPB_E_CUDA (-2)    /* CUDA runtime error (message has the details) */
#define PPB_E_NOMEM (-3)
#define PPB_E_STATE (-4)   /* call not valid in the batch's current state */

/* arithmetic of a batch */
#define PPB_F32 0
#define PPB_F64 1

/* per-world outcome (ppb_get_status); a world whose status is not PPB_ST_OK is
 * halted at the step the reference would have raised in.  Same numbering as
 * oracle/billiards_oracle.c. */
#define PPB_ST_OK 0
#define PPB_ST_ASSERT_TCOL_NEG 1 /* primitives.py:610 */
#define PPB_ST_ASSERT_OVERLAP 2  /* geometry.py:413  */
#define PPB_ST_ASSERT_THETA 3    /* geometry.py:436-438 */
#define PPB_ST_ASSERT_THETA_A 4  /* geometry.py:447 */
#define PPB_ST_ASSERT_INTPOINT 5 /* dynamics.py:34 */
#define PPB_ST_TRAP_AMISS 6      /* dynamics.py:43-45 */
#define PPB_ST_MATH_DOMAIN 7     /* math.acos / math.asin ValueError, geometry.py:174,444 */
#define PPB_ST_ASSERT_TCOL_EQ 8  /* primitives.py:664 */
#define PPB_ST_ITER_LIMIT 9      /* engine guard, see ppb_config.max_substeps */

#define PPB_MAX_BALLS 32 /* ball slots per world */
#define PPB_MAX_WALLS 16 /* wall slots per world (64 collision edges) */

typedef struct ppb_batch ppb_batch; /* opaque */

typedef struct ppb_config {
  int32_t n_worlds;
  int32_t n_balls;        /* ball slots per world, 1..PPB_MAX_BALLS */
  int32_t n_walls;        /* wall slots per world, 0..PPB_MAX_WALLS */
  int32_t dtype;          /* PPB_F32 | PPB_F64 */
  int32_t canvas_w;       /* World(xSz, ySz), primitives.py:857 */
  int32_t canvas_h;
  int32_t device;         /* CUDA device ordinal */
  int32_t max_substeps;  /* guard on the sub-step loop of one step() (the reference would spin for
                             minutes or never return); 0 = default 1048576.  A world that exceeds it is
                             halted with PPB_ST_ITER_LIMIT */
  double dt;              /* Dynamics(deltaT=0.01), primitives.py:536 */
  double gx, gy;          /* Dynamics.g_ (a Point), primitives.py:541, 566 */
  double friction;        /* extension, 0 = reference behaviour: at the start of every step all velocities are
                             scaled by max(0, 1 - friction*dt) (semi-implicit Euler); cached collision times are
                             rescaled by the inverse, so the event-driven step stays exact */
  int64_t event_capacity; /* records kept for ppb_read_events; 0 = count only */
  const int32_t *order;   /* n_walls + n_balls entries: object insertion order, wall q -> q,
                             ball b -> n_walls + b; NULL = walls first, then balls */
} ppb_config;

/* one collision event = one call of Dynamics.resolve_collision (primitives.py:659) */
typedef struct ppb_event {
  int32_t world;
  int32_t seq;     /* running number of the event inside its world (call order) */
  int32_t step;    /* number of completed step() calls before this one */
  int16_t ball;    /* ball slot */
  int16_t partner; /* object index as in `order` (wall q -> q, ball b -> n_walls + b) */
} ppb_event;

/* drawing attributes of one ball slot (BallDef, primitives.py:362-371) */
typedef struct ppb_ball_style {
  int32_t s_thick;    /* sThick; the sprite is 2*(radius+sThick) pixels wide, primitives.py:39 */
  float fill[4];      /* fColor r,g,b,a */
  float stroke[4];    /* sColor r,g,b,a */
} ppb_ball_style;

/* drawing attributes of one wall slot */
typedef struct ppb_wall_style {
  int32_t kind;       /* 0 = GenericWall (primitives.py:221), 1 = Wall from WallDef (primitives.py:266) */
  float fill[4];      /* fColor r,g,b,a */
} ppb_wall_style;

PPB_API const char *ppb_last_error(void);
PPB_API int ppb_version(void);

/* pm.World(...) + pm.Dynamics(world, g, deltaT) for n_worlds worlds
 * (primitives.py:536-563, 857-877).  Allocates the device state. */
PPB_API int ppb_create(const ppb_config *cfg, ppb_batch **out);
PPB_API int ppb_destroy(ppb_batch *b);

/* world.add_object(wall) for every wall slot of worlds [world0, world0+count):
 * `wall` is [count][n_walls][4][2] vertices in Bbox order (primitives.py:224-228,
 * 291-300; geometry.py:471-505). */
PPB_API int ppb_set_walls(ppb_batch *b, int32_t world0, int32_t count, const double *wall, void *stream);

/* world.add_object(ball) + Dynamics._include_object for every ball slot
 * (primitives.py:556-563, 916-929): pos, vel are [count][n_balls][2]; rad, mass
 * [count][n_balls] (rad < 0 = empty slot).  Resets the collision caches of these
 * worlds (tCol_ = 0, objCol_ = None, all balls queued) and their status. */
PPB_API int ppb_set_balls(ppb_batch *b, int32_t world0, int32_t count, const double *pos, const double *vel,
                  const double *rad, const double *mass, void *stream);

/* Dynamics.apply_force(name, force, forceT) for every ball of worlds [world0, world0+count)
 * (primitives.py:580-594): `force` is [count][n_balls][2]; vel += ((0.5*forceT)*forceT) * ((1.0/mass)*force),
 * i.e. the reference's displacement formula used as a velocity increment (quirk Q7), evaluated in fp64 in the
 * reference's operation order.  Collision caches are NOT refreshed, exactly as in the reference: call it before the
 * first step (DataSaver does, dataio.py:382-416) or follow it with ppb_requeue.  Empty slots are skipped. */
PPB_API int ppb_apply_force(ppb_batch *b, int32_t world0, int32_t count, const double *force, double force_t,
                    void *stream);

/* parameters of the random rectangular tables of DataSaver(isRect=True) (dataio.py:21-27) */
typedef struct ppb_sampler_params {
  double arena;            /* arenaSz */
  double w_thick;          /* wThick */
  double mn_wlen, mx_wlen; /* mnWLen, mxWLen */
  double mn_ball, mx_ball; /* mnBallSz, mxBallSz */
  double mn_force, mx_force;
  double force_t;          /* forceT of the initial kick (1.0 in dataio.py:414) */
  uint64_t seed;
  int64_t world_index0;    /* global index of world 0 of this batch (sharded jobs) */
  int32_t max_tries;       /* give up on a world after this many rejections (dataio.py:359-362: 500) */
  int32_t n_theta;         /* 0 = rectangular table (isRect); 1..4 = rhombus, side angle drawn from theta[] (wTheta) */
  double theta[4];         /* degrees, dataio.py:243-306 */
} ppb_sampler_params;

/* DataSaver._generate_model + apply_force (dataio.py:198-416) for worlds [world0, world0+count),
 * drawn ON THE DEVICE: rectangular or rhombus cage of 4 walls (the batch needs n_walls >= 4), n_balls balls placed by
 * rejection sampling, initial velocities from random kicks; collision caches reset as by
 * ppb_set_balls.  Distribution-equivalent to the reference (same loop structure and fp64
 * arithmetic) on a Philox stream keyed by (seed, world_index0 + world); the stream-exact sampler is
 * the host-side DataSaver.  *n_failed receives the number of worlds that could not be filled
 * (left empty: every ball slot has radius -1). */
PPB_API int ppb_sample_rect_worlds(ppb_batch *b, int32_t world0, int32_t count, const ppb_sampler_params *params,
                           int32_t *n_failed, void *stream);

/* ball.set_position / ball.set_velocity without touching the caches
 * (primitives.py:462-466); NULL leaves that quantity unchanged. */
PPB_API int ppb_update_balls(ppb_batch *b, int32_t world0, int32_t count, const double *pos, const double *vel,
                     void *stream);

/* Dynamics.add_all_dynamic_collision_queue (primitives.py:697-700): every ball of
 * these worlds is rescanned at the start of the next step. */
PPB_API int ppb_requeue(ppb_batch *b, int32_t world0, int32_t count, void *stream);

/* ball.get_position / get_velocity (primitives.py:450-457): [count][n_balls][2] each, may be NULL */
PPB_API int ppb_get_balls(ppb_batch *b, int32_t world0, int32_t count, double *pos, double *vel, void *stream);

/* ball.get_radius / get_mass (primitives.py:468-475): [count][n_balls] each, may be NULL; radius < 0 = empty slot */
PPB_API int ppb_get_radius_mass(ppb_batch *b, int32_t world0, int32_t count, double *rad, double *mass, void *stream);

/* wall vertices back: [count][n_walls][4][2] (wall.bbox_.vert_, geometry.py:471-505) */
PPB_API int ppb_get_walls(ppb_batch *b, int32_t world0, int32_t count, double *wall, void *stream);

/* Dynamics.tCol_ / objCol_ (primitives.py:548-549): [count][n_balls]; partner -1 = (None, None) */
PPB_API int ppb_get_collision_cache(ppb_batch *b, int32_t world0, int32_t count, double *tcol, int32_t *partner,
                            void *stream);

/* per-world outcome and the number of completed steps ([count] each, may be NULL);
 * for a halted world steps_done is the step it raised in */
PPB_API int ppb_get_status(ppb_batch *b, int32_t world0, int32_t count, int32_t *status, int32_t *steps_done,
                   void *stream);

/* Drawing attributes (BallDef / WallDef / GenericWall colours).  Builds the
 * pre-rendered ball sprites (get_ball_im, primitives.py:33-61) on the device for
 * every integer radius in [r_min, r_max]. */
PPB_API int ppb_set_styles(ppb_batch *b, const ppb_ball_style *balls /*[n_balls]*/,
                   const ppb_wall_style *walls /*[n_walls]*/, int32_t r_min, int32_t r_max, void *stream);

/* get_ball_im (primitives.py:33-61): copies the pre-rendered sprite of ball slot `slot` at integer
 * radius `radius` to `out_host`: [sz][sz][4] uint8 with sz = 2*(radius + s_thick), premultiplied
 * alpha, channel order Color.r, Color.g, Color.b, alpha.  *sz_out receives sz. */
PPB_API int ppb_read_sprite(ppb_batch *b, int32_t slot, int32_t radius, uint8_t *out_host, int64_t cap_bytes,
                    int32_t *sz_out, void *stream);

/* Dynamics.set_g (primitives.py:566): replaces the gravity vector; collision caches are left alone
 * exactly as the reference does. */
PPB_API int ppb_set_gravity(ppb_batch *b, double gx, double gy);

/* n_steps x Dynamics.step() (primitives.py:628-656) for every world, asynchronous.
 *   traj_dev   : NULL or [n_steps][n_worlds][n_balls][2] in the batch dtype -- the
 *                positions after each step (what dataio.py:123-128 records)
 *   frames_dev : NULL or [n_steps / render_every][n_worlds][H][W][4] uint8 --
 *                World.generate_image() after every render_every-th step
 *                (primitives.py:991-1008, dataio.py:119-120)                       */
PPB_API int ppb_step_async(ppb_batch *b, int32_t n_steps, void *traj_dev, uint8_t *frames_dev,
                   int32_t render_every, void *stream);
/* Scheduling note: when more than one frame is requested, one advance launch covers a chunk of up
 * to 64 steps and leaves a position snapshot per drawn step; the chunk's render launches run on a
 * stream owned by the batch and read those snapshots while the next chunk's advance launch may
 * already start; `stream` waits for the last render before the call's work counts as complete,
 * so stream-ordered consumers of `frames_dev` need no extra synchronisation. */

/* Drains the collision queue without stepping: Dynamics.time_to_collide_all (primitives.py:720-733)
 * for every ball of every world whose queue is non-empty (after ppb_set_balls / ppb_requeue), i.e.
 * dynamics.get_toc_ball_wall / get_toc_ball_ball (dynamics.py:8-136) against every other object.
 * Positions, velocities and step counters are unchanged; results are read with
 * ppb_get_collision_cache / ppb_get_collision_response.  Asynchronous. */
PPB_API int ppb_scan_async(ppb_batch *b, void *stream);

/* Dynamics.nrmlCol_ (primitives.py:550; for a ball partner the contact tangent, dynamics.py:95) and
 * Ball.futureVel_ (dynamics.py:129): [count][n_balls][2] each, may be NULL */
PPB_API int ppb_get_collision_response(ppb_batch *b, int32_t world0, int32_t count, double *nrml, double *future_vel,
                               void *stream);

/* As ppb_step_async, with the frames written round-robin into `ring_slots` frame slots
 * ([ring_slots][n_worlds][H][W][4] uint8): the frame of the j-th rendered step goes to slot
 * j % ring_slots.  Lets a long run keep its frames in a bounded buffer (a consumer drains it). */
PPB_API int ppb_step_ring_async(ppb_batch *b, int32_t n_steps, void *traj_dev, uint8_t *frames_dev, int32_t ring_slots,
                        int32_t render_every, void *stream);

/* World.generate_image() of the current state: [n_worlds][H][W][4] uint8, asynchronous.
 * Every call that draws frames (this one, ppb_step_async / ppb_step_ring_async with frames_dev, ppb_simulate_host*
 * with frames_host) fails with PPB_E_STATE when some occupied ball slot has a non-integer radius or one outside the
 * [r_min, r_max] given to ppb_set_styles: the reference builds a ball's sprite with the radius as an array shape
 * (primitives.py:39-40), so there is no sprite such a ball could be drawn from.  The radii are measured on the device
 * the first time frames are requested after ppb_set_balls / ppb_sample_rect_worlds (one short synchronisation). */
PPB_API int ppb_render_async(ppb_batch *b, uint8_t *frames_dev, void *stream);

/* Host-buffer convenience used by the reference-facing API (DataSaver.fetch,
 * dataio.py:131-196): runs n_steps with the copies inside the call.
 *   traj_host   : NULL or [n_steps][n_worlds][n_balls][2] float32 (dataio.py:103-104 casts to fp32)
 *   frames_host : NULL or [n_steps / render_every][n_worlds][H][W][4] uint8
 * Host buffers should be page-locked for full copy bandwidth. */
PPB_API int ppb_simulate_host(ppb_batch *b, int32_t n_steps, float *traj_host, uint8_t *frames_host,
                      int32_t render_every, void *stream);

/* Frames reach the host run-length packed: a rendered canvas is opaque and mostly flat colour, so ppb_simulate_host*
 * send rgb | run_length << 24 tokens over PCIe (about a seventh of the bytes of a 64x64 billiards frame) and a pool of
 * host threads rebuilds the RGBA pixels in `frames_host`; the bytes delivered are exactly the device frames.  Frames
 * that are not opaque or do not compress travel raw.
 *
 * ppb_download_frames: the same path for any device array of `n_frames` canvases ([n_frames][H][W][4] uint8, the
 * batch's canvas size; e.g. the frames of ppb_step_async or a ring) -> `frames_host`; synchronous. */
PPB_API int ppb_download_frames(ppb_batch *b, const uint8_t *frames_dev, int64_t n_frames, uint8_t *frames_host, void *stream);

/* The same without waiting: returns once the steps and the device -> host copies are enqueued (the copies run on a
 * stream owned by the batch, behind the kernels).  The host buffers of a call are complete after
 * ppb_host_sync(b, 0, stream) -- or after ppb_host_sync(b, 1, stream) once a later call has been enqueued
 * (keep_in_flight = 1 waits for every call but the most recent one).  Until then they must not be read, and the
 * next call must be given different host buffers.  Lets a caller upload the inputs of the next batch
 * (ppb_set_walls / ppb_set_balls on `stream`) while the frames of this one are still crossing PCIe -- the two
 * directions do not share bandwidth. */
PPB_API int ppb_simulate_host_async(ppb_batch *b, int32_t n_steps, float *traj_host, uint8_t *frames_host,
                            int32_t render_every, void *stream);
PPB_API int ppb_host_sync(ppb_batch *b, int32_t keep_in_flight, void *stream);

/* The ball-centred crops of DataSaver.fetch(cropSz) (dataio.py:160-182) for every frame, world and ball:
 *   frames_dev : [n_frames][n_worlds][H][W][4] uint8 (what ppb_step_async rendered)
 *   traj_dev   : [n_frames][n_worlds][n_balls][2] in the batch dtype -- the ball centres in those frames
 *   crops_dev  : [n_frames][n_worlds][n_balls][crop_sz][crop_sz][3] uint8, white where the crop leaves the canvas
 *   pos_dev    : NULL or [n_frames][n_worlds][n_balls][2] float32 -- the centres in crop coordinates (dataio.py:178-179)
 * mode 0 = DataSaver.fetch (offsets rounded, positions re-based); mode 1 = CustomWorld.get_glimpse (dataio.py:505-518:
 * offsets truncated with int(), pos_dev = the centres unchanged).  Asynchronous. */
PPB_API int ppb_crop_async(ppb_batch *b, const uint8_t *frames_dev, const void *traj_dev, int32_t n_frames, int32_t crop_sz,
                   int32_t mode, uint8_t *crops_dev, float *pos_dev, void *stream);

/* The procSz step of DataSaver.fetch(cropSz, procSz) (dataio.py:183-189): scipy.misc.imresize(imBall, (procSz, procSz)),
 * i.e. PIL's bilinear resize, of every crop ppb_crop_async produced, and the rescaling that goes with it.
 *   crops_dev : [n_frames][n_worlds][n_balls][crop_sz][crop_sz][3] uint8 (crop_sz <= 255)
 *   out_dev   : [n_frames][n_worlds][n_balls][proc_sz][proc_sz][3] uint8
 *   pos_dev   : NULL or the positions ppb_crop_async returned, multiplied in place by proc_sz / crop_sz
 *   vel_dev   : NULL or [n_frames][n_worlds][n_balls][2] float32; row i-1 receives (centre_i - centre_{i-1}) * proc_sz / crop_sz
 *               from traj_dev ([n_frames][n_worlds][n_balls][2], batch dtype), the last row is left untouched
 * The resampler restates Pillow's 8-bit path (libImaging/Resample.c).  Asynchronous. */
PPB_API int ppb_resize_crops_async(ppb_batch *b, const uint8_t *crops_dev, const void *traj_dev, int32_t n_frames, int32_t crop_sz,
                           int32_t proc_sz, uint8_t *out_dev, float *pos_dev, float *vel_dev, void *stream);

/* DynamicsHorizon.get_data (primitives.py:839-852) for every frame of a recorded trajectory in one launch:
 *   traj_dev : [n_rows][n_worlds][n_balls][2] positions in the batch dtype; row 0 = before the first step
 *   out_dev  : [n_rows - lookahead][n_worlds][n_balls][2 * lookahead] float32;
 *              out[t][w][b][2k + c] = float32(traj[t + k + 1][w][b][c] - traj[t + k][w][b][c])
 * Asynchronous. */
PPB_API int ppb_horizon_async(ppb_batch *b, const void *traj_dev, int32_t n_rows, int32_t lookahead, float *out_dev, void *stream);

/* utils.detect_collision's per-frame test (utils.py:38-49) on a recorded trajectory:
 *   traj_dev  : [n_steps][n_items][2] float32 positions (n_items = worlds x balls)
 *   flags_dev : [n_items][n_steps - 2] uint8, 1 where |diff(diff(pos))|^2 >= thresh * |diff(pos)|^2
 * No batch is needed; runs on the current device.  Asynchronous. */
PPB_API int ppb_collision_flags_async(const float *traj_dev, int64_t n_items, int32_t n_steps, float thresh,
                              uint8_t *flags_dev, void *stream);

/* events recorded since the last ppb_set_balls on world 0 / ppb_clear_events:
 * copies up to `cap` records (unordered across worlds; sort by (world, seq)) and
 * returns the total number produced in *n_total (may exceed event_capacity). */
PPB_API int ppb_read_events(ppb_batch *b, ppb_event *out, int64_t cap, int64_t *n_total, void *stream);
PPB_API int ppb_clear_events(ppb_batch *b, void *stream);

/* counters since creation: [0] kernels launched by this batch, [1] world-steps that ran the
 * event-driven loop of Dynamics.step (all other world-steps were event-free: move_object only),
 * [2] sub-step loop iterations inside that loop, [3] collision events, [4] world rescans (one rescan =
 * time_to_collide_all for every ball of a world = n_balls x (4 n_walls + n_balls - 1) pair tests),
 * [5..7] reserved (0) */
PPB_API int ppb_get_counters(ppb_batch *b, int64_t out[8], void *stream);

/* raw device pointers for zero-copy consumers (e.g. torch via __cuda_array_interface__):
 * which = 0: positions   [n_worlds][n_balls] x {x, y} (batch dtype)
 *         1: tCol        [n_worlds][n_balls]
 *         2: walls       [n_worlds][n_walls][4][2]
 *         3: radius/mass [n_worlds][n_balls] x {r, m}
 *         4: velocities  [n_worlds][n_balls] x {vx, vy}
 * State written through these pointers takes effect after ppb_requeue (collision caches; the renderer's copy of the
 * radii); walls written this way also need ppb_set_styles again (the renderer's per-wall cache).                  */
PPB_API int ppb_device_ptr(ppb_batch *b, int32_t which, void **ptr, int64_t *nbytes);

/* time (ms) and launch count of the integrate / collide / render kernels, measured with
 * CUDA events recorded around every launch on the launching stream and summed over all
 * stepping calls since ppb_set_profiling(b, mode) / the previous ppb_get_kernel_times (which
 * waits for the recorded events and resets the accumulation).
 * mode: 0 off, 1 every kernel, 2 the render kernel only, 3 integrate + collide only.  An event pair between
 * two kernels of a stream keeps the second from being scheduled while the first drains (a few us per
 * launch), so mode 2 is what a throughput measurement that also wants the dominant kernel's time uses. */
PPB_API int ppb_set_profiling(ppb_batch *b, int32_t on);
PPB_API int ppb_get_kernel_times(ppb_batch *b, float ms[3], int32_t launches[3]);

#ifdef __cplusplus
}
#endif
#endif /* PYPHY_B200_H */
