/* worldb200 -- C-ABI of the B200-native WORLD vocoder hot path.
 *
 * Drop-in boundary for yukara-ikemiya/world-class.  The reference has no FFI; its
 * boundary is the public section of its C++ class headers.  Every entry point below
 * replaces one reference interface, cited as file:line relative to /root/reference.
 * The source-compatible C++ class headers in this directory (harvest.hpp,
 * cheaptrick.hpp, d4c.hpp, synthesis.hpp, codec.hpp) are inline shims over this ABI.
 *
 * Conventions
 *   - plain C: opaque handles, POD option structs, raw pointers and sizes; no CUDA or
 *     torch types (a stream is passed as void*, NULL = the library's own stream).
 *   - every function that can fail returns an int status (WB_OK = 0); the reference
 *     returns void and has undefined behaviour on bad input.
 *   - `*_compute`      : HOST pointers, identical argument meaning to the reference's
 *                        compute(); synchronous; copies H2D/D2H internally.
 *   - `*_compute_dev`  : DEVICE pointers, 2-D outputs are contiguous [frames][bins];
 *                        asynchronous on the given stream.
 *   - the randn() stream is process-global like the reference's
 *     (src/world_matlabfunctions.cpp:243-264) and is consumed in the same order and
 *     amount, so a fresh process running Harvest->CheapTrick->D4C->Synthesis draws the
 *     same noise as the reference's serial build.
 */
#ifndef WORLDB200_H
#define WORLDB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK 0
#define WB_ERR_CUDA 1
#define WB_ERR_ARG 2
#define WB_ERR_UNSUPPORTED 3

/* include/harvest.hpp:16-28 (HarvestOption; defaults src/harvest.cpp:52-56) */
typedef struct WbHarvestOption {
  double f0_floor;            /* 71.0  */
  double f0_ceil;             /* 800.0 */
  double frame_period;        /* 5.0 ms */
  double target_fs;           /* 8000.0 */
  double channels_in_octave;  /* 40.0 */
  int use_cos_table;          /* 0; the approximate window table is not supported (exact path only) */
} WbHarvestOption;

/* include/cheaptrick.hpp:14-20 (defaults src/cheaptrick.cpp:22-24) */
typedef struct WbCheapTrickOption {
  double q1;        /* -0.15 */
  double f0_floor;  /* 71.0 */
  int fft_size;     /* 0 = derive from fs and f0_floor */
} WbCheapTrickOption;

/* include/d4c.hpp:16-20 (default src/d4c.cpp:31-33) */
typedef struct WbD4COption {
  double threshold; /* 0.85 */
} WbD4COption;

typedef struct wb_harvest wb_harvest_t;
typedef struct wb_cheaptrick wb_cheaptrick_t;
typedef struct wb_d4c wb_d4c_t;
typedef struct wb_synthesis wb_synthesis_t;
typedef struct wb_pipeline wb_pipeline_t;
typedef struct wb_synthesis_stream wb_synthesis_stream_t;

/* ---- library ---------------------------------------------------------------------- */
int wb_init(int device);                /* optional; otherwise lazily on first use (current device) */
const char *wb_version(void);
int wb_device_synchronize(void);

void wb_harvest_option_default(WbHarvestOption *opt);      /* src/harvest.cpp:52-56 */
void wb_cheaptrick_option_default(WbCheapTrickOption *opt); /* src/cheaptrick.cpp:22-24 */
void wb_d4c_option_default(WbD4COption *opt);               /* src/d4c.cpp:31-33 */

/* ---- randn() stream (src/world_matlabfunctions.cpp:243-264) ------------------------- */
int wb_randn_reseed(void);                            /* back to the reference's initial state */
int wb_randn_get_state(unsigned int state[4]);        /* x, y, z, w */
int wb_randn_set_state(const unsigned int state[4]);
int wb_randn_skip(unsigned long long n_calls);        /* as if randn() had been called n times */
int wb_randn_fill(double *out, int n);                /* HOST out; next n values; advances the stream */

/* ---- stand-alone FFT with the reference wrapper's semantics -------------------------
 * include/world_fft.hpp:33-41 + src/world_fft.cpp:31-77: forward = e^{+i}, backward = e^{-i},
 * unnormalised; complex numbers are interleaved (re, im) doubles; HOST pointers;
 * n = power of two in [128, 16384] (c2c: up to 8192); `batch` independent transforms back to back. */
int wb_fft_r2c(const double *in, int n, int batch, double *out);   /* fft_plan_dft_r2c_1d + fft_execute */
int wb_fft_c2r(const double *in, int n, int batch, double *out);   /* fft_plan_dft_c2r_1d + fft_execute */
int wb_fft_c2c(const double *in, int n, int batch, int sign, double *out); /* sign: 1 = FFT_FORWARD, 2 = FFT_BACKWARD */

/* ---- Harvest (include/harvest.hpp:31-44) ------------------------------------------------ */
int wb_harvest_get_samples(int fs, int x_length, double frame_period);          /* src/harvest.cpp:173-181 */
int wb_harvest_create(int fs, const WbHarvestOption *opt_or_null, wb_harvest_t **out); /* src/harvest.cpp:69-103 */
void wb_harvest_destroy(wb_harvest_t *h);
/* src/harvest.cpp:183-208; temporal_positions and f0 hold wb_harvest_get_samples() entries */
/* One call analyses up to about 16 minutes of audio (1024 overlap-save blocks at the 8 kHz analysis rate; its
 * scratch grows with the length, about 32 MB per second of audio).  Longer inputs return WB_ERR_UNSUPPORTED before anything is
 * launched; process them in segments (worldb200.parallel).  The reference accepts any length. */
int wb_harvest_compute(wb_harvest_t *h, const double *x, int x_length, double *temporal_positions, double *f0);
int wb_harvest_compute_dev(wb_harvest_t *h, const double *d_x, int x_length, double *d_temporal_positions,
                           double *d_f0, void *stream);
/* test hook: copies n_bytes of a named internal device buffer of the last compute() to `out` */
int wb_harvest_debug_read(wb_harvest_t *h, const char *name, void *out, unsigned long long n_bytes);

/* ---- CheapTrick (include/cheaptrick.hpp:23-38) ------------------------------------- */
int wb_cheaptrick_get_fft_size(int fs, double f0_floor);       /* src/cheaptrick.cpp:97-100 */
double wb_cheaptrick_get_f0_floor(int fs, int fft_size);       /* src/cheaptrick.cpp:102-105 */
int wb_cheaptrick_create(int fs, const WbCheapTrickOption *opt_or_null, wb_cheaptrick_t **out); /* :27-45 */
void wb_cheaptrick_destroy(wb_cheaptrick_t *h);
int wb_cheaptrick_fft_size(const wb_cheaptrick_t *h);
/* src/cheaptrick.cpp:48-95; spectrogram = f0_length separately allocated rows of fft_size/2+1 */
int wb_cheaptrick_compute(wb_cheaptrick_t *h, const double *x, int x_length,
                          const double *temporal_positions, const double *f0, int f0_length,
                          double **spectrogram);
int wb_cheaptrick_compute_dev(wb_cheaptrick_t *h, const double *d_x, int x_length,
                              const double *d_temporal_positions, const double *d_f0, int f0_length,
                              double *d_spectrogram, void *stream);

/* ---- D4C (include/d4c.hpp:23-36) ---------------------------------------------------- */
int wb_get_number_of_aperiodicities(int fs);                    /* src/d4c.cpp:65-67, src/codec.cpp:211-214 */
int wb_d4c_create(int fs, const WbD4COption *opt_or_null, wb_d4c_t **out);   /* src/d4c.cpp:35-44 */
void wb_d4c_destroy(wb_d4c_t *h);
/* src/d4c.cpp:113-173; aperiodicity = f0_length separately allocated rows of fft_size/2+1 */
int wb_d4c_compute(wb_d4c_t *h, const double *x, int x_length, const double *temporal_positions,
                   const double *f0, int f0_length, int fft_size, double **aperiodicity);
int wb_d4c_compute_dev(wb_d4c_t *h, const double *d_x, int x_length, const double *d_temporal_positions,
                       const double *d_f0, int f0_length, int fft_size, double *d_aperiodicity, void *stream);

/* ---- Synthesis (include/synthesis.hpp:29-51) ------------------------------------------ */
int wb_synthesis_create(int fs, int fft_size, double frame_period_ms, wb_synthesis_t **out); /* src/synthesis.cpp:30-56 */
void wb_synthesis_destroy(wb_synthesis_t *h);
/* src/synthesis.cpp:77-177; `out` (out_length samples) is fully overwritten */
int wb_synthesis_compute(wb_synthesis_t *h, const double *f0, int f0_length,
                         const double *const *spectrogram, const double *const *aperiodicity,
                         int out_length, double *out);
/* f0_upper_bound: host-known upper bound of max(f0) (sizes the pulse buffers without a
 * device->host round trip); <= 0 = unknown (one stream synchronisation inside the call). */
int wb_synthesis_compute_dev(wb_synthesis_t *h, const double *d_f0, int f0_length,
                             const double *d_spectrogram, const double *d_aperiodicity,
                             int out_length, double *d_out, double f0_upper_bound, void *stream);

/* ---- streaming Synthesis (the real-time use the demo mentions, test/test.cpp:353-358; the reference itself
 * only has the one-shot compute(), src/synthesis.cpp:77-177) ---------------------------------------------------
 * Frames are pushed in pieces of any size; each call returns the samples that became final (no later pulse can
 * reach them).  The phase sum, the pulse list and the randn() position are carried between calls, so the
 * concatenation of everything returned equals wb_synthesis_compute() on all frames bit for bit.  HOST pointers;
 * spectrogram / aperiodicity are contiguous [n_frames][fft_size/2+1].  At most out_capacity samples are written
 * per call (what does not fit comes with the next call).  finish(): no more frames; out_length_total as in
 * compute(); call it until *n_out == 0.  The process-global randn() stream belongs to the stream object from
 * the first push to the last finish.  f0_upper_bound: a bound on the f0 values (sizes the pulse buffers;
 * <= 0 = 1000 Hz). */
int wb_synthesis_stream_create(int fs, int fft_size, double frame_period_ms, double f0_upper_bound,
                               wb_synthesis_stream_t **out);
void wb_synthesis_stream_destroy(wb_synthesis_stream_t *s);
int wb_synthesis_stream_push(wb_synthesis_stream_t *s, const double *f0, const double *spectrogram,
                             const double *aperiodicity, int n_frames, double *out, int out_capacity, int *n_out);
int wb_synthesis_stream_finish(wb_synthesis_stream_t *s, int out_length_total, double *out, int out_capacity, int *n_out);

/* ---- codec (include/codec.hpp:23-88; src/codec.cpp:211-325) ------------------------------
 * Same argument meaning as the reference's free functions; rows are separately allocated. */
int wb_code_aperiodicity(const double *const *aperiodicity, int f0_length, int fs, int fft_size,
                         double **coded_aperiodicity);                       /* src/codec.cpp:216-235 */
int wb_decode_aperiodicity(const double *const *coded_aperiodicity, int f0_length, int fs, int fft_size,
                           double **aperiodicity);                           /* src/codec.cpp:237-265 */
int wb_code_spectral_envelope(const double *const *spectrogram, int f0_length, int fs, int fft_size,
                              int number_of_dimensions, double **coded_spectral_envelope); /* :267-296 */
int wb_decode_spectral_envelope(const double *const *coded_spectral_envelope, int f0_length, int fs,
                                int fft_size, int number_of_dimensions, double **spectrogram); /* :298-325 */
/* contiguous DEVICE arrays; kind: 0 code ap, 1 decode ap, 2 code sp, 3 decode sp; asynchronous */
int wb_codec_dev(int kind, const double *d_in, int f0_length, int fs, int fft_size, int number_of_dimensions,
                 double *d_out, void *stream);

/* ---- parameter modification between analysis and synthesis (test/test.cpp:201-243) --------------
 * F0 scaling (f0[i] *= f0_shift; pass NaN to leave F0 alone) and spectral stretching by `ratio` (each row:
 * log, interp1 from the axis ratio*i/fft_size*fs to i/fft_size*fs, exp; for ratio < 1 the bins from
 * int(fft_size/2*ratio) on repeat the bin before them; pass ratio <= 0 to leave the spectrogram alone).
 * In place, like the reference demo.  *_dev: contiguous DEVICE arrays, asynchronous on `stream`. */
int wb_parameter_modification(double *f0, int f0_length, double **spectrogram, int fs, int fft_size,
                              double f0_shift, double ratio);
int wb_parameter_modification_dev(double *d_f0, int f0_length, double *d_spectrogram, int fs, int fft_size,
                                  double f0_shift, double ratio, void *stream);

/* ---- whole chain, device resident -----------------------------------------------------
 * The call sequence of test/test.cpp:288-384 (Harvest -> CheapTrick -> D4C -> Synthesis) with
 * every intermediate kept in HBM.  NULL option pointers = defaults; NULL output pointers in
 * *_run_dev = internal buffers.  spectrogram / aperiodicity are contiguous [f0_length][fft_size/2+1]. */
int wb_pipeline_create(int fs, const WbHarvestOption *hopt, const WbCheapTrickOption *copt,
                       const WbD4COption *dopt, wb_pipeline_t **out);
void wb_pipeline_destroy(wb_pipeline_t *p);
/* fresh != 0: every run restarts its own randn() stream at the reference's seed (one utterance ==
 * one reference process); such pipelines may run concurrently on different streams (batches). */
int wb_pipeline_set_fresh_rng(wb_pipeline_t *p, int fresh);
/* use_graph != 0: repeated wb_pipeline_run_dev calls with identical arguments replay a captured CUDA
 * graph (one launch for the whole chain).  The internal buffers are used unless all outputs are given. */
int wb_pipeline_set_graph(wb_pipeline_t *p, int use_graph);
/* apply the parameter modification above between analysis and synthesis (the demo's call order,
 * test/test.cpp:318-332): the returned f0 / spectrogram are the modified ones.  NaN / <= 0 switch a part off. */
int wb_pipeline_set_modification(wb_pipeline_t *p, double f0_shift, double ratio);
int wb_pipeline_fft_size(const wb_pipeline_t *p);
int wb_pipeline_f0_length(const wb_pipeline_t *p, int x_length);      /* src/harvest.cpp:173-181 */
int wb_pipeline_out_length(const wb_pipeline_t *p, int x_length);     /* test/test.cpp:362-363 */
int wb_pipeline_run_dev(wb_pipeline_t *p, const double *d_x, int x_length, double *d_temporal_positions,
                        double *d_f0, double *d_spectrogram, double *d_aperiodicity, double *d_y,
                        int y_length, void *stream);
int wb_pipeline_run(wb_pipeline_t *p, const double *x, int x_length, double *temporal_positions_or_null,
                    double *f0_or_null, double *spectrogram_or_null, double *aperiodicity_or_null,
                    double *y, int y_length);

/* wav in -> wav out, the demo's whole flow (test/test.cpp:288-384 between wavread, tools/audioio.cpp:228-253,
 * and wavwrite, :121-184): HOST 16-bit PCM in both directions; the two sample-format conversions run on the
 * device, so 2 bytes per sample cross PCIe instead of 8. */
int wb_pipeline_run_pcm16(wb_pipeline_t *p, const short *pcm_in, int x_length, short *pcm_out, int y_length);
/* like wb_pipeline_run with HOST outputs narrowed to fp32 on the device (contiguous matrices; NULL = not
 * wanted).  The arithmetic is fp64 throughout; only the delivered arrays are rounded (to nearest). */
int wb_pipeline_run_f32(wb_pipeline_t *p, const double *x, int x_length, float *f0_or_null, float *spectrogram_or_null,
                        float *aperiodicity_or_null, float *y, int y_length);

/* ---- one long stream sharded over ranks (BASELINE configs[3]: frames shard, one exchange step) -----------
 * CheapTrick (src/cheaptrick.cpp:48-95), D4C (src/d4c.cpp:113-173) and Synthesis (src/synthesis.cpp:77-177) of a
 * RANGE of a stream whose whole f0 contour is known (gathered from the ranks' Harvest runs).  A frame's
 * position in the process-global randn() stream (src/world_matlabfunctions.cpp:243-264) and the sequential
 * phase sum that places the pulses (src/synthesis.cpp:257-264) are recomputed for the whole stream on every
 * rank -- they are cheap functions of f0 -- so the rows / samples of a range are bit-identical to those of an
 * unsharded run.  Device pointers; asynchronous on `stream`; d_f0_all / d_ap0_all hold f0_length entries;
 * *_rows address the first row of the range ([rows][fft_size/2+1], contiguous).  Per rank, in this order:
 *   begin; envelope (writes d_ap0_all[frame_begin, frame_end)); [all-gather d_ap0_all]; aperiodicity; synthesis; end
 * (envelope / aperiodicity / synthesis may be repeated for several ranges in between).
 * The sp / ap rows given to synthesis must cover every frame a pulse reaching into the sample range
 * interpolates between: frames floor((sample_begin - fft_size) / fs / frame_period) .. ceil((sample_end +
 * fft_size) / fs / frame_period), clipped to the stream. */
int wb_pipeline_stream_begin_dev(wb_pipeline_t *p, const double *d_f0_all, int f0_length, int out_length, void *stream);
/* The same with the sample range [sample_begin, sample_end) this rank will synthesise (the union of the ranges of its
 * later wb_pipeline_stream_synthesis_dev calls): the exact phase sum still runs over the whole stream, but the
 * per-sample passes that turn it into a pulse list are restricted to that range (two thirds of the replicated work).
 * sample_end < 0 = the whole stream. */
int wb_pipeline_stream_begin_range_dev(wb_pipeline_t *p, const double *d_f0_all, int f0_length, int out_length,
                                       int sample_begin, int sample_end, void *stream);
int wb_pipeline_stream_envelope_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                    int f0_length, int frame_begin, int frame_end, double *d_sp_rows,
                                    double *d_ap0_all, void *stream);
/* the envelope call in two halves: Love Train first (so that the exchange of its decisions overlaps CheapTrick) */
int wb_pipeline_stream_lovetrain_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                     int f0_length, int frame_begin, int frame_end, double *d_ap0_all, void *stream);
int wb_pipeline_stream_cheaptrick_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                      int f0_length, int frame_begin, int frame_end, double *d_sp_rows, void *stream);
int wb_pipeline_stream_aperiodicity_dev(wb_pipeline_t *p, const double *d_x, int x_length, const double *d_f0_all,
                                        const double *d_ap0_all, int f0_length, int frame_begin, int frame_end,
                                        double *d_ap_rows, void *stream);
int wb_pipeline_stream_synthesis_dev(wb_pipeline_t *p, int f0_length, const double *d_sp_rows, const double *d_ap_rows,
                                     int row_begin, int n_rows, int out_length, int sample_begin, int sample_end,
                                     double *d_out, void *stream);
/* after the last range: leaves the randn() state where one reference process would have left it */
int wb_pipeline_stream_end_dev(wb_pipeline_t *p, void *stream);

/* test / bench hook: copies n_bytes of a named internal device buffer of the last run to `out` */
int wb_pipeline_debug_read(wb_pipeline_t *p, const char *name, void *out, unsigned long long n_bytes);

/* ---- data formats either side of the path (N2, N3) --------------------------------
 * Sample formats of the reference's wav reader / writer as DEVICE conversions, asynchronous on `stream`:
 * x = pcm / 32768 (tools/audioio.cpp:232-249 at 16 bit), pcm = clamp(trunc(x * 32767)) (:176-180). */
int wb_pcm16_to_f64_dev(const short *d_pcm, int n, double *d_x, void *stream);
int wb_f64_to_pcm16_dev(const double *d_x, int n, short *d_pcm, void *stream);
int wb_f64_to_f32_dev(const double *d_in, unsigned long long n, float *d_out, void *stream);
/* The reference's parameter files (tools/parameterio.hpp:24-117, tools/parameterio.cpp:60-244), byte-compatible
 * both ways; same argument meaning, but failures are status codes instead of a printf.  HOST code. */
int wb_write_f0(const char *filename, int f0_length, double frame_period, const double *temporal_positions,
                const double *f0, int text_flag);                                   /* parameterio.cpp:60-90 */
int wb_read_f0(const char *filename, double *temporal_positions, double *f0);       /* :92-119 */
double wb_get_header_information(const char *filename, const char *parameter);      /* :121-147 */
int wb_write_spectral_envelope(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                               int number_of_dimensions, const double *const *spectrogram);   /* :149-177 */
int wb_read_spectral_envelope(const char *filename, double **spectrogram);          /* :179-199 */
int wb_write_aperiodicity(const char *filename, int fs, int f0_length, double frame_period, int fft_size,
                          int number_of_dimensions, const double *const *aperiodicity);       /* :201-223 */
int wb_read_aperiodicity(const char *filename, double **aperiodicity);              /* :225-244 */
/* the same containers from / to one contiguous matrix (kind 0 = "SPEC", 1 = "AP  "; row_stride in doubles) */
int wb_write_parameter_matrix(int kind, const char *filename, int fs, int f0_length, double frame_period,
                              int fft_size, int number_of_dimensions, const double *matrix, long long row_stride);
int wb_read_parameter_matrix(int kind, const char *filename, double *matrix, long long row_stride);
/* The reference's RIFF reader / writer (tools/audioio.hpp, tools/audioio.cpp): mono PCM only */
int wb_wavwrite(const double *x, int x_length, int fs, int nbit, const char *filename);  /* audioio.cpp:121-184 */
int wb_get_audio_length(const char *filename);        /* :186-226; 0 = cannot open, -1 = header not accepted */
int wb_wavread(const char *filename, int *fs, int *nbit, double *x);                      /* :228-253 */
int wb_wavread_pcm16(const char *filename, int *fs, short *pcm);  /* the raw 16-bit samples (for *_pcm16) */

/* ---- measurement hooks (bench.py) ------------------------------------------------------ */
unsigned long long wb_launch_count(void);  /* kernels launched by this library so far */
/* decimate() of /root/reference/include/world_matlabfunctions.hpp:81 (src/world_matlabfunctions.cpp:184-210) on the
 * GPU: zero-phase third-order IIR low-pass, every r-th sample; r = 2 .. 12.  HOST pointers; y receives
 * wb_decimate_length(x_length, r) samples. */
int wb_decimate_length(int x_length, int r);
int wb_decimate(const double *x, int x_length, int r, double *y);
/* Errors detected on the device (more pulses than the f0 bound allows for, a smoothing width beyond the scratch
 * capacity) are flagged in a per-handle word.  Host-pointer entry points return it themselves; after asynchronous
 * *_dev calls ask here: waits for `stream` (NULL = the library's), returns WB_OK or the flagged status, clears it. */
int wb_harvest_last_error(wb_harvest_t *h, void *stream);
int wb_cheaptrick_last_error(wb_cheaptrick_t *h, void *stream);
int wb_d4c_last_error(wb_d4c_t *h, void *stream);
int wb_synthesis_last_error(wb_synthesis_t *h, void *stream);
int wb_pipeline_last_error(wb_pipeline_t *p, void *stream);
/* sharded streams with an f0 contour that does not come from Harvest: upper bound of its values (sizes the pulse
 * buffers); <= 0 = Harvest's f0_ceil * 1.25 */
int wb_pipeline_set_stream_f0_bound(wb_pipeline_t *p, double f0_upper_bound);
int wb_measure_fp64_peak(double *tflops);  /* dependent-free DFMA chains on every SM: the fp64 roofline of this GPU */
void *wb_stream(void);                     /* the library's own cudaStream_t */
void wb_profile_enable(int on);            /* bracket every kernel launch with CUDA events */
void wb_profile_reset(void);
int wb_profile_collect(void);              /* synchronises; folds pending events into totals */
int wb_profile_query(const char *kernel_name, double *total_ms, int *count);
int wb_profile_names(char *buf, int buf_len);  /* ';'-separated kernel names seen so far */

#ifdef __cplusplus
}
#endif
#endif /* WORLDB200_H */
