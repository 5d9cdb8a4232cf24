/*
 * innfer_b200 -- C ABI of the B200-native RRDB/ESRGAN inference path.
 *
 * The reference (victorca25/iNNfer) is pure Python and has no FFI; its boundary for this path is
 * the object protocol between run.py:Model and architectures.get_network.
 * Every entry point below names the reference interface it replaces (file:line relative to the
 * reference tree).  The Python shim in innfer_b200/ keeps the reference's classes and calls these
 * functions through ctypes.
 *
 * Conventions: all functions return 0 on success or a negative INNFER_E_* code and never throw;
 * innfer_last_error() returns a thread-local message for the last failure.  Pointers marked
 * "device" are CUDA device pointers on the handle's device; "host" pointers are ordinary host
 * memory.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 * stream ordered and the functions do not synchronise unless stated.  A handle must be used from
 * one host thread at a time.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef INNFER_B200_H_
#define INNFER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INNFER_OK 0
#define INNFER_E_INVALID (-1)      /* bad argument / shape */
#define INNFER_E_UNSUPPORTED (-2)  /* valid for the reference, not built here (fails loudly) */
#define INNFER_E_CUDA (-3)         /* CUDA runtime / driver error */
#define INNFER_E_STATE (-4)        /* call order (e.g. forward before finalize, missing weights) */
#define INNFER_E_NOMEM (-5)

/* element types of tensors crossing the ABI */
#define INNFER_F16 0
#define INNFER_F32 1
#define INNFER_U8 2

typedef struct innfer_rrdb innfer_rrdb; /* opaque network handle */

/* Constructor kwargs of RRDBNet (architectures/RRDBNet_arch.py:17-19) as produced by
 * get_network_G_config (utils/defaults.py:20-44).  gc is accepted for completeness; like the
 * reference (RRDBNet_arch.py:26) the blocks are always built with 32 growth channels. */
typedef struct innfer_rrdb_cfg {
  int32_t in_nc;
  int32_t out_nc;
  int32_t nf;
  int32_t nb;
  int32_t gc;
  int32_t scale; /* upscale: 1, 2, 3, 4, 8 */
  int32_t plus;  /* ESRGAN+ residual paths: conv1x1 + extra adds (RRDBNet_arch.py:129,155-160) */
  int32_t fp16;  /* 1: fp16 storage + tcgen05 fp16 MMA with fp32 accumulate; 0: fp32 mode */
} innfer_rrdb_cfg;

/* Constructor kwargs of SRResNet (architectures/SRResNet_arch.py:16-17) as produced by
 * get_network_G_config (utils/defaults.py:53-67): no norm, ReLU, mode CNA.  SURVEY.md 8(f) rank 1. */
typedef struct innfer_srresnet_cfg {
  int32_t in_nc;
  int32_t out_nc;
  int32_t nf;            /* 32 or 64 */
  int32_t nb;            /* ResNetBlocks */
  int32_t scale;         /* 1, 2, 3, 4, 8 */
  int32_t upsample_mode; /* 0: pixelshuffle (reference default, block.py:333-346), 1: upconv */
  float res_scale;       /* ResNetBlock residual scaling */
  int32_t fp16;
} innfer_srresnet_cfg;

/* Constructor kwargs of PPON (architectures/PPON_arch.py:17-18) as produced by get_network_G_config
 * (utils/defaults.py:68-77): nf = 64 (RRBlock_32 is hard-wired to 64 channels), LeakyReLU.  SURVEY.md 8(f) rank 3. */
typedef struct innfer_ppon_cfg {
  int32_t in_nc;
  int32_t out_nc;
  int32_t nf;            /* 64 */
  int32_t nb;            /* RRBlock_32 blocks of the content module (24) */
  int32_t scale;         /* 1, 2, 3, 4, 8 */
  float alpha;           /* out_p = alpha * PRM(..) + out_s */
  int32_t fp16;
} innfer_ppon_cfg;

/* Constructor kwargs of PAN (architectures/PAN_arch.py:104-106) as produced by get_network_G_config
 * (utils/defaults.py:78-89); ups_inter_mode is 'nearest'.  SURVEY.md 8(f) rank 3. */
typedef struct innfer_pan_cfg {
  int32_t in_nc;
  int32_t out_nc;          /* == in_nc (the bilinear skip adds the input image to the output) */
  int32_t nf;              /* 40; multiple of 8, at most 64 */
  int32_t unf;             /* 24; channels of the upsampling stages (nf when scale == 1) */
  int32_t nb;              /* SCPA blocks per trunk (16) */
  int32_t scale;           /* 1, 2, 3, 4, 8 */
  int32_t self_attention;  /* FSA block (max-pooled self attention) after the trunk */
  int32_t double_scpa;     /* second SCPA trunk + trunk_conv2 */
  int32_t fp16;
} innfer_pan_cfg;

/* Constructor kwargs of UnetGenerator (architectures/UNet_arch.py:21-22) / ResnetGenerator
 * (architectures/ResNet_arch.py:20-22) as produced by get_network_G_config (utils/defaults.py:99-135), plus the one piece
 * of module state that changes the arithmetic: run.py:297 keeps pix2pix in training mode (meval False), where
 * BatchNorm2d normalises with the statistics of the batch.  upsample_mode 'deconv', padding 'reflect', no dropout.
 * SURVEY.md 8(f) rank 4. */
typedef struct innfer_i2i_cfg {
  int32_t kind;   /* 0: UnetGenerator (pix2pix), 1: ResnetGenerator (CycleGAN) */
  int32_t in_nc;
  int32_t out_nc;
  int32_t ngf;    /* multiple of 8 */
  int32_t depth;  /* num_downs (5..12) or n_blocks */
  int32_t norm;   /* 0: BatchNorm2d, 1: InstanceNorm2d (then the convolutions carry a bias) */
  int32_t train;  /* BatchNorm2d: 1 = batch statistics (module.training), 0 = running statistics */
  int32_t fp16;
  int32_t unit_io; /* 1 (ResnetGenerator only): images in [0, 1] at both ends -- the [-1, 1] normalisation run.py wraps
                    * around these networks (np2tensor(normalize) / tensor2np(denormalize), utils.py:136-161,
                    * run.py:420,430) folded into the first conv (doubled weights, bias - sum of weights: exact with
                    * reflection padding) and the last layer ((tanh + 1) / 2), so that innfer_rrdb_upscale_u8 /
                    * chop_forward_ex serve the CLI loop on uint8 frames like they do for the SR networks */
} innfer_i2i_cfg;

typedef struct innfer_tile {
  int32_t y0, x0; /* low-res origin of the tile */
} innfer_tile;

/* ---- library ------------------------------------------------------------------------------- */
const char* innfer_last_error(void);
const char* innfer_version(void);
/* number of innfer CUDA kernels launched by this process so far (bench.py "gpu_launches") */
uint64_t innfer_kernel_launches(void);

/* ---- network handle: replaces architectures.get_network + nn.Module.load_state_dict/.to(device)
 *      (architectures/__init__.py:5-40, run.py:90-101) ------------------------------------------- */
int innfer_rrdb_create(const innfer_rrdb_cfg* cfg, int device, innfer_rrdb** out);
/* SRResNet handle; every other innfer_rrdb_* call (load / finalize / forward / chop_forward / upscale_u8 /
 * tile-range / destroy) works on it unchanged.  Keys: "model.0", "model.1.sub.<i>.res.0|2", "model.1.sub.<nb>",
 * "model.2|5" (pixel-shuffle convs, 4*nf filters), "model.8", "model.10" (4x). */
int innfer_srresnet_create(const innfer_srresnet_cfg* cfg, int device, innfer_rrdb** out);
/* PPON handle (same calls as above work on it); the forward result is the third output of PPON.forward, out_p,
 * which is what run.py uses (run.py:191-192,220-221).  Keys: "CFEM.0", "CFEM.1.sub.<i>.RB<r>.{c1,d1..d8,c2}",
 * "CFEM.1.sub.<nb>", "SFEM|PFEM.<0|1>.RB<r>.*", "CRM|SRM|PRM.<1,4,6,8>" (4x). */
int innfer_ppon_create(const innfer_ppon_cfg* cfg, int device, innfer_rrdb** out);
/* PAN handle (same calls as above work on it).  Keys: "conv_first", "SCPA_trunk.<i>.{conv1_a,conv1_b,k1.0,
 * PACnv.k2,PACnv.k3,PACnv.k4,conv3}", "trunk_conv", "FSA.{gamma,conv_f,conv_g,conv_h}", "upsample.<1,2.conv,4,..>",
 * "conv_last"; with several upsampling stages the reference drops the LeakyReLU after HRconv (block.py:204-207
 * flattens through children(), which yields the shared activation instance once) and so does this path. */
int innfer_pan_create(const innfer_pan_cfg* cfg, int device, innfer_rrdb** out);
/* pix2pix UNet / CycleGAN ResNet generator handle (scale 1; the same calls as above work on it).  Keys: UNet
 * "model.model.0", "model.model.1.model.{1,2,3.model...,5,6}", ..., "model.model.3"; ResNet "model.1|2", "model.4|5",
 * "model.7|8", "model.<10+b>.conv_block.{1,2,5,6}", "model.<10+nb>|<11+nb>", "model.<13+nb>|<14+nb>", "model.<17+nb>".
 * innfer_rrdb_forward treats its batch as ONE reference call (train-mode BatchNorm statistics over the batch);
 * chop_forward treats every tile as its own call, as run.py:187-197 does. */
int innfer_i2i_create(const innfer_i2i_cfg* cfg, int device, innfer_rrdb** out);
/* one state-dict entry under its REFERENCE key name ("model.0.weight",
 * "model.1.sub.3.RDB2.conv4.0.bias", "model.1.sub.23.weight", "model.10.bias", ...); host fp32. */
int innfer_rrdb_load(innfer_rrdb* h, const char* key, const float* host_data, const int64_t* shape,
                     int ndim);
/* verifies that every parameter arrived (strict load), repacks OIHW -> kernel layout, uploads. */
int innfer_rrdb_finalize(innfer_rrdb* h);
void innfer_rrdb_destroy(innfer_rrdb* h);
/* upper bound of tiles pushed through the trunk per batch (workspace memory vs. launch count);
 * default 95 (two batches per 1080p frame, ~20 GB of workspace at 4x) */
int innfer_rrdb_set_max_batch(innfer_rrdb* h, int max_tiles);

/* device-side timing of the conv sequence (CUDA events on the launching stream around every
 * per-batch trunk+tail pass): enable/reset with innfer_rrdb_profile(h, 1); innfer_rrdb_profile_read
 * synchronises on the recorded events and returns the summed milliseconds and the number of conv
 * kernels launched inside them.  Used by bench.py for the roofline line; no reference equivalent. */
int innfer_rrdb_profile(innfer_rrdb* h, int enable);
int innfer_rrdb_profile_read(innfer_rrdb* h, double* conv_ms, uint64_t* conv_launches);
/* innfer_rrdb_profile(h, 2): events around EVERY conv launch, accumulated per kernel family (kernel, Cin->Cout,
 * flags).  The events defeat the programmatic-dependent-launch overlap of consecutive convs, so these are the times
 * of isolated launches; mode 1 times the real schedule.  innfer_rrdb_profile_families writes one line per family,
 * "name\tlaunches\ttotal_ms\talgorithmic_flop\talgorithmic_bytes\n", into buf (if cap suffices); *needed = bytes
 * required including the terminating NUL. */
int innfer_rrdb_profile_families(innfer_rrdb* h, char* buf, uint64_t cap, uint64_t* needed);

/* ---- forward: replaces RRDBNet.forward on one batch (RRDBNet_arch.py:50-62) -------------------
 * x: device NCHW [n][in_nc][h][w], y: device NCHW [n][out_nc][scale*h][scale*w]; dtype INNFER_F16
 * or INNFER_F32 for both. */
int innfer_rrdb_forward(innfer_rrdb* h, const void* x, int n, int hgt, int wid, void* y, int dtype,
                        void* stream);

/* ---- chop_forward: replaces Model.chop_forward = extract_patches_2d -> per-tile forward ->
 *      recompose_tensor (run.py:167-202, utils/utils.py:318-445) -------------------------------
 * x: device NCHW [1][in_nc][H][W]; y: device NCHW [1][out_nc][scale*H][scale*W]. */
int innfer_rrdb_chop_forward(innfer_rrdb* h, const void* x, int H, int W, int patch_size,
                             double step, void* y, int dtype, void* stream);

/* same with independent element types at the two ends: x INNFER_F16/F32 (NCHW) or INNFER_U8 (HWC BGR, np2tensor
 * fused), y INNFER_F16/F32 (NCHW) or INNFER_U8 (HWC BGR, tensor2np fused).  This is what lets a model chain
 * (run.py:424-426: `for mod in models: t_out = mod(t_out)`) stay on the device: uint8 frame -> fp16 tensor ->
 * ... -> uint8 frame, with float tensors between the models exactly as in the reference. */
int innfer_rrdb_chop_forward_ex(innfer_rrdb* h, const void* x, int x_dtype, int H, int W, int patch_size,
                                double step, void* y, int y_dtype, void* stream);

/* ---- image in, image out: np2tensor -> chop_forward -> tensor2np fused (run.py:421-431,
 *      utils/utils.py:164-248).  img: HOST uint8 HWC BGR [H][W][3]; out: HOST uint8 HWC BGR
 *      [scale*H][scale*W][3].  Copies run on `stream`; the call returns after the result landed
 *      in `out` (it synchronises the stream). */
int innfer_rrdb_upscale_u8(innfer_rrdb* h, const uint8_t* img, int H, int W, int patch_size,
                           double step, uint8_t* out, void* stream);
/* same with DEVICE uint8 buffers and no synchronisation */
int innfer_rrdb_upscale_u8_device(innfer_rrdb* h, const uint8_t* img, int H, int W, int patch_size,
                                  double step, uint8_t* out, void* stream);

/* ---- tile-sharded execution across GPUs (the reference is single-GPU and runs the
 *      tile loop of run.py:187-197 serially).  One process per GPU; the frame's owner exposes its
 *      tile buffer and its low-res frame through CUDA IPC, every rank computes a contiguous range
 *      of the row-major tile list and its last conv stores the finished tiles straight into the
 *      owner's buffer (peer stores over NVLink, no NCCL, no staging copy); the owner then blends.
 *      Results are bit-identical for any number of ranks. ------------------------------------- */
/* owner: (re)allocate the handle's tile buffer for an HxW frame; returns its device pointer */
int innfer_rrdb_tile_buffer(innfer_rrdb* h, int H, int W, int patch_size, double step, void** ptr,
                            uint64_t* bytes, uint64_t* bytes_per_tile);
/* size of the tile buffer of an HxW frame (what a rank must allocate, with innfer_device_alloc, to own frames) */
int innfer_rrdb_tile_bytes(innfer_rrdb* h, int H, int W, int patch_size, double step, uint64_t* bytes_total,
                           uint64_t* bytes_per_tile, int* ntiles);
/* any rank: tiles [t_begin, t_end) of the frame `img` (device or peer pointer; INNFER_U8 HWC BGR or
 * NCHW F16/F32) -> tiles_base + t * bytes_per_tile (device or peer pointer) */
int innfer_rrdb_forward_tile_range(innfer_rrdb* h, const void* img, int img_dtype, int H, int W,
                                   int patch_size, double step, int t_begin, int t_end,
                                   void* tiles_base, void* stream);
/* owner: recompose_tensor (+ tensor2np for INNFER_U8) over a complete tile buffer */
int innfer_rrdb_blend_tiles(innfer_rrdb* h, const void* tiles_base, int H, int W, int patch_size,
                            double step, void* dst, int dst_dtype, void* stream);
/* CUDA IPC plumbing for the two calls above (handles are 64 opaque bytes) */
int innfer_ipc_export(const void* device_ptr, uint8_t handle[64]);
int innfer_ipc_open(const uint8_t handle[64], void** device_ptr);
int innfer_ipc_close(void* device_ptr);
/* plain cudaMalloc / cudaFree on `device` (IPC needs allocations that are not sub-allocated by a
 * caching allocator) */
int innfer_device_alloc(int device, uint64_t bytes, void** ptr);
int innfer_device_free(void* ptr);
int innfer_device_memset(void* ptr, int value, uint64_t bytes); /* synchronous */
/* cudaMemcpyAsync(cudaMemcpyDefault) on `stream`: pinned host <-> device and device <-> peer-mapped device copies */
int innfer_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream);
/* Stream-ordered cross-GPU flags (csrc/sync_ops.cu): `flags` is a HOST array of n (<= 32) device pointers to
 * uint32 counters (local or peer-mapped).  signal: after all earlier work of `stream`, store `value` into each
 * (system-scope release).  wait: hold `stream` until every counter has reached `value`; if that takes longer than
 * timeout_ms the kernel increments *err_flag (device uint32, may be NULL) and lets the stream continue, so a
 * protocol bug cannot hang the GPU.  These replace host barriers in the tile-sharded mode. */
int innfer_stream_signal(void* const* flags, int n, uint32_t value, void* stream);
int innfer_stream_wait(void* const* flags, int n, uint32_t value, void* err_flag, uint64_t timeout_ms, void* stream);

/* Debugging aid (not part of the reference-facing surface): device buffer of 3072 int64 that receives
 * clock64 samples of CTA 0 of the row-streaming conv kernel selected by INNFER_TRACE_NCH; NULL disables. */
int innfer_debug_set_trace(void* device_buf);
/* Measurement aid: `iters` back-to-back launches (after `warm` untimed ones) of one 3x3 conv on a wide batch
 * of B random H x W images; *ms_out = device time of the timed launches. */
int innfer_debug_conv_loop(int Cin, int Cout, int B, int H, int W, int with_res, int warm, int iters, float* ms_out);
/* host -> device copy into such an allocation (synchronises `stream`) */
int innfer_device_upload(void* device_dst, const void* host_src, uint64_t bytes, void* stream);

/* ---- tiling geometry: replaces the index arithmetic of extract_patches_2d
 *      (utils/utils.py:349-365).  Writes up to `cap` tiles in row-major order; *n = tile count,
 *      *tile_size = min(H, W, patch_size). */
int innfer_tiles_plan(int H, int W, int patch_size, double step, innfer_tile* out, int cap, int* n,
                      int* tile_size);

/* ---- standalone operators (used by the parity tests, same kernels as the network) ----------- */
/* image -> tiles: np2tensor + extract_patches_2d.  src device: NCHW fp16/fp32 [1][C][H][W] or
 * uint8 HWC BGR; dst device planar-chunk tiles [ntiles][ceil(C/16)*2][p][p][8] fp16. */
int innfer_image_to_tiles(const void* src, int src_dtype, int C, int H, int W, int patch_size,
                          double step, void* dst_tiles, void* stream);
/* recompose_tensor (+ tensor2np when dst_dtype is INNFER_U8).  tiles: device planar-chunk
 * [ntiles][1][P][P][8] fp16 with P = scale*min(H,W,patch_size). */
int innfer_blend(const void* tiles, int H, int W, int patch_size, double step, int scale, int C,
                 void* dst, int dst_dtype, void* stream);
/* same for fp32 tiles [ntiles][1][P][P][8] float (the -no_fp16 mode keeps fp32 through the blend) */
int innfer_blend_f32(const float* tiles, int H, int W, int patch_size, double step, int scale, int C, void* dst,
                     int dst_dtype, void* stream);
/* one fused conv_block (architectures/block.py:213-254) [+ nearest Upsample in front,
 * block.py:348-361] [+ LeakyReLU] [+ alpha1*. + res1] on NCHW device tensors; weights host fp32
 * OIHW.  Builds, runs and frees a temporary layer -- a test/bring-up entry point, not a hot path.
 * x [n][Cin][h][w], res1 (or NULL) and y [n][Cout][up*h][up*w] all of `dtype`.
 * use_fp32_kernel: 0 = fp16 kernels on the tiled layout, 1 = fp32 direct kernel, 2 = fp16 kernels on the wide
 * batch layout the engine uses (row-streaming kernel for Cout = 32).  In mode 2, res1 == x (the same pointer, Cin >= Cout)
 * means "the residual is the first Cout channels of the conv's own input", the shape conv5 has inside a dense block
 * (RRDBNet_arch.py:164-165: x5 * 0.2 + x). */
int innfer_conv3x3(const void* x, int n, int Cin, int hgt, int wid, const float* w_oihw,
                   const float* bias, int Cout, int up, int lrelu, const void* res1, float alpha1,
                   void* y, int dtype, int use_fp32_kernel, void* stream);

/* one layer of the image-to-image generators on NCHW device tensors, through the same kernels as the networks
 * (csrc/i2i.cu): nn.Conv2d / nn.ConvTranspose2d (UNet_arch.py:108-137, ResNet_arch.py:52-88) with kernel k, stride 1|2,
 * zero or reflection padding (ReflectionPad2d(pad) in front of an unpadded conv), optional bias; then optionally
 * norm: 1 InstanceNorm2d, 2 BatchNorm2d with batch statistics (affine norm_weight / norm_bias, null = 1 / 0); then
 * act: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 tanh.  final_path 1 runs bias + act in the conv epilogue (the route of a
 * network's last layer), 0 through the fp32 scratch + normalisation kernels.  w is host fp32 OIHW (IOHW for transposed),
 * x [n][Cin][h][w], y [n][Cout][h'][w'] device tensors of `dtype` (F16: tcgen05 kernel, F32: direct kernel).
 * Synchronises the stream -- a test/bring-up entry point, not a hot path. */
int innfer_gen_conv(const void* x, int n, int Cin, int hgt, int wid, const float* w, const float* bias, int Cout, int k,
                    int stride, int pad, int transposed, int out_pad, int reflect, int norm, const float* norm_weight,
                    const float* norm_bias, int act, int final_path, void* y, int dtype, void* stream);

/* launches of the halo-tile variant of the generator convolution by this process (the parity tests use it to know
 * which of the two tensor-core kernels served a layer; INNFER_I2I_HALO=0|2 in the environment forces the choice) */
uint64_t innfer_debug_i2i_halo_launches(void);
/* forwards of the image-to-image generators served by replaying a recorded CUDA graph (small shapes: the launch sequence of
 * a (buffers, shape) combination is recorded the second time it is seen; INNFER_I2I_GRAPH=0 disables, 2 records every size) */
uint64_t innfer_debug_i2i_graph_replays(void);

/* ---- -cf colour correction: replaces color_fix (utils/utils.py:278-315) with srgb2linear /
 *      linear2srgb (utils/colors.py:29-60), cv2.resize(INTER_CUBIC) and cv2.GaussianBlur((3,3),0).
 *      lr: device uint8 [h][w][3]; sr: device uint8 [H][W][3]; out: device uint8 [H][W][3].
 *      scratch is allocated internally and cached per thread. */
int innfer_color_fix(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W, uint8_t* out,
                     void* stream);
/* host-buffer convenience wrapper (H2D, kernels, D2H, synchronises) */
int innfer_color_fix_host(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W,
                          uint8_t* out, int device);

#ifdef __cplusplus
}
#endif
#endif /* INNFER_B200_H_ */
