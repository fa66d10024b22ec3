"""Sharding of the WORLD hot path over the GPUs of one box (one process per GPU).

Two cases:

* batches of independent utterances (BASELINE configs[2]): round-robin over ranks, no data-path
  collective;
* one long stream (BASELINE configs[3]): the stream is cut into segments that overlap by a halo,
  every rank analyses / re-synthesises its segments as independent utterances, the halos are
  discarded and ONE all-gather stitches the segment cores back into the output stream.  Every
  segment reproduces the reference run on that segment (fresh randn() stream per segment); the
  halo (default 1 s >> the longest analysis window of 3/40 s and Harvest's 300 ms contour
  padding) makes the cores independent of where the cuts fall.

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes
import numpy as np


HARVESTS_IN_FLIGHT = 3   # Harvest segments of a rank analysed concurrently (one Harvest object + stream each)


def shard_indices(n_items, rank, world):
    """Round-robin assignment of item indices to `rank`."""
    return list(range(rank, n_items, world))


def plan_segments(n_samples, fs, segment_seconds=30.0, halo_seconds=1.0):
    """Cuts [0, n_samples) into segment cores of `segment_seconds`; returns a list of dicts
    core=(a, b) and padded=(a - halo, b + halo) clipped to the stream, sample indices."""
    seg = max(1, int(round(segment_seconds * fs)))
    halo = max(0, int(round(halo_seconds * fs)))
    out = []
    a = 0
    while a < n_samples:
        b = min(n_samples, a + seg)
        out.append({"core": (a, b), "padded": (max(0, a - halo), min(n_samples, b + halo))})
        a = b
    return out


def extract_core(y_padded, seg):
    """Drops the halos of a re-synthesised padded segment (lengths may differ by the frame grid:
    the synthesised length of a segment is floor-aligned to the frame period)."""
    (a, b), (pa, _pb) = seg["core"], seg["padded"]
    core = np.zeros(b - a, dtype=np.float64)
    src = y_padded[a - pa:a - pa + (b - a)]
    core[:len(src)] = src
    return core


def process_stream(x, fs, process_segment, segment_seconds=30.0, halo_seconds=1.0, group=None):
    """Shards a long stream over the ranks of `group` and returns the stitched output on every rank.

    process_segment(x_padded) -> y_padded must map a padded segment to its re-synthesis (same
    sample grid).  Works without torch.distributed being initialised (single process).
    """
    try:
        import torch
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    except Exception:  # pragma: no cover
        distributed = False
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    segs = plan_segments(len(x), fs, segment_seconds, halo_seconds)
    mine = shard_indices(len(segs), rank, world)
    seg_len = max(s["core"][1] - s["core"][0] for s in segs)
    per_rank = (len(segs) + world - 1) // world
    # local cores in a fixed-size [per_rank, seg_len] block so that one all-gather suffices
    local = np.zeros((per_rank, seg_len), dtype=np.float64)
    for slot, k in enumerate(mine):
        s = segs[k]
        y_pad = process_segment(x[s["padded"][0]:s["padded"][1]])
        core = extract_core(np.asarray(y_pad, dtype=np.float64), s)
        local[slot, :len(core)] = core
    if distributed:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t_local = torch.from_numpy(local).to(dev)
        gathered = torch.empty((world,) + t_local.shape, dtype=t_local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered.view(-1), t_local.view(-1), group=group) if hasattr(
            dist, "all_gather_into_tensor") and backend == "nccl" else dist.all_gather(
            [gathered[r] for r in range(world)], t_local, group=group)
        blocks = gathered.cpu().numpy()
    else:
        blocks = local[None]
    y = np.zeros(len(x), dtype=np.float64)
    for k, s in enumerate(segs):
        r, slot = k % world, k // world
        a, b = s["core"]
        y[a:b] = blocks[r, slot, :b - a]
    return y


def process_batch(items, process_item, group=None):
    """Round-robin shards a list of independent utterances; returns {index: result} of the local
    shard (no collective on the data path; gather the results with `gather_objects` if wanted)."""
    try:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    except Exception:  # pragma: no cover
        distributed = False
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    return {i: process_item(items[i]) for i in shard_indices(len(items), rank, world)}


# ---- exact sharding of one long stream (BASELINE configs[3]) -----------------------
#
# process_stream() above treats segments as independent utterances (fresh randn() stream each): simple, but
# the stitched result is not what ONE reference process would produce for the whole stream.  The functions
# below shard the same work so that it IS: every rank
#   1. runs Harvest on its share of the stream (sub-segments with a halo, aligned so that the decimation phase
#      and the 1 ms analysis grid coincide with those of the whole stream) and ONE all-gather assembles the
#      whole f0 contour;
#   2. recomputes the cheap sequential parts for the whole stream (randn() positions = prefix sums of the
#      per-frame draw counts, the phase sum that places the pulses) and computes CheapTrick / Love Train for
#      its own frame rows only;
#   3. a second small all-gather exchanges the Love Train decisions (they decide how many randn() draws the
#      D4C body of every EARLIER frame consumed), then D4C and Synthesis run on the rank's rows / samples with
#      whole-stream randn positions -- bit-identical to an unsharded run given the same f0;
#   4. one all-gather stitches the waveform.
# The collectives are torch.distributed's (NCCL on GPUs, gloo in the CPU tests); the compute is the C-ABI's
# wb_pipeline_stream_* entry points.


def bind_to_gpu_node(device_index):
    """One process per GPU: keep this process -- and the threads it creates from here on (the library's copy pool) --
    on the CPUs NVML reports as local to the GPU's NUMA node, so page-locked buffers allocated afterwards and the
    host side of every H2D / D2H copy stay off the inter-socket link.  Call before allocating pinned memory.
    Returns a dict describing what was done (`bound` False when NVML or the affinity call is unavailable, or the
    node's CPUs are outside this process's cpuset); never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        prop = torch.cuda.get_device_properties(device_index)
        bus_id = "%08x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinityWithinScope(handle, (n_cpu + 63) // 64, pynvml.NVML_AFFINITY_SCOPE_NODE)
        local = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        info.update(pci=bus_id, node_cpus=len(local), allowed_cpus=len(allowed))
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus), first_cpu=cpus[0], last_cpu=cpus[-1])
    except Exception as e:   # noqa: BLE001 (a missing NVML / an odd cpuset must not stop the job)
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


def decimation_ratio(fs, target_fs=8000.0):
    """Harvest's decimation ratio (src/harvest.cpp:81-82)."""
    r = int(fs / target_fs + 0.5)          # matlab_round of a positive number
    return max(1, min(12, r))


class StreamPlan:
    """Pure planning (no GPU): who owns which frames / samples of a stream of n_samples at fs, and which
    padded Harvest segments reproduce the whole-stream analysis grid.

    Rank boundaries and segment cuts fall on whole seconds, so that with 1000 / frame_period integer every
    segment's local frame j is the stream's frame j + offset at exactly the same time, and the padded end of
    a segment is congruent to n_samples modulo the decimation ratio, which gives the segment the
    whole-stream decimation phase (decimate() starts at x_length % r - 1, world_matlabfunctions.cpp:201-206)."""

    def __init__(self, n_samples, fs, world, frame_period=5.0, fft_size=2048, segment_seconds=120, halo_seconds=2,
                 target_fs=8000.0):
        fps = 1000.0 / frame_period
        if abs(fps - round(fps)) > 1e-9:
            raise ValueError("exact stream sharding needs 1000 / frame_period to be an integer")
        self.n, self.fs, self.world, self.fp, self.fft_size = int(n_samples), int(fs), int(world), float(frame_period), int(fft_size)
        self.fps = int(round(fps))
        self.r = decimation_ratio(fs, target_fs)
        if self.fs % self.r:
            raise ValueError("exact stream sharding needs fs to be a multiple of the decimation ratio")
        self.f0_length = int(1000.0 * self.n / self.fs / self.fp) + 1                    # harvest.cpp:173-176
        self.out_length = int((self.f0_length - 1) * self.fp / 1000.0 * self.fs) + 1     # test/test.cpp:362-363
        seconds = -(-self.n // self.fs)                                                   # ceil
        seg, halo = max(1, int(segment_seconds)), max(1, int(halo_seconds))
        # rank r owns seconds [cut[r], cut[r + 1])
        self.cut = [(seconds * k) // self.world for k in range(self.world + 1)]
        self.segments = []      # per rank: list of dict(core frames, padded samples, frame offset)
        self.frames = []        # per rank: (frame_begin, frame_end) of the rows it owns
        for k in range(self.world):
            s0, s1 = self.cut[k], self.cut[k + 1]
            fb = min(self.f0_length, s0 * self.fps)
            fe = self.f0_length if k == self.world - 1 else min(self.f0_length, s1 * self.fps)
            self.frames.append((fb, fe))
            segs = []
            # the rank's seconds in the fewest pieces of at most `seg` seconds, evenly long (every piece pays its halo)
            n_seg = max(1, -(-(s1 - s0) // seg))
            step = max(1, -(-(s1 - s0) // n_seg))
            a = s0
            while a < s1:
                b = min(s1, a + step)
                pa = max(0, a - halo) * self.fs
                pb = min(self.n, (b + halo) * self.fs)
                pb -= (pb - self.n) % self.r                  # whole-stream decimation phase
                cfb = a * self.fps
                cfe = self.f0_length if (k == self.world - 1 and b == s1) else min(self.f0_length, b * self.fps)
                if cfe > cfb:
                    segs.append({"padded": (pa, pb), "frames": (cfb, cfe), "frame_offset": (pa // self.fs) * self.fps})
                a = b
            self.segments.append(segs)
        # samples: any partition works; cut where the rank's first frame starts
        cuts = [0] + [min(self.out_length, int(self.frames[k][0] * self.fp / 1000.0 * self.fs)) for k in range(1, self.world)] + [self.out_length]
        self.samples = [(cuts[k], cuts[k + 1]) for k in range(self.world)]
        # rows (frames) a rank needs for its samples: every frame a pulse reaching into them interpolates between
        self.rows = []
        for k in range(self.world):
            sa, sb = self.samples[k]
            fb, fe = self.frames[k]
            lo = int(max(0, sa - self.fft_size) / self.fs * self.fps) - 1
            hi = int((sb + self.fft_size) / self.fs * self.fps) + 3
            self.rows.append((max(0, min(fb, lo)), min(self.f0_length, max(fe, hi))))


def gather_ranges(local, ranges, total, group=None, device=None):
    """Assembles a 1-D float64 array of `total` entries from per-rank contiguous pieces (`local` is this
    rank's piece, ranges[k] = (begin, end) of rank k) with ONE all-gather.  torch tensors in / out."""
    import torch
    import torch.distributed as dist
    world = len(ranges)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        assert len(local) == total
        return local
    rank = dist.get_rank(group)
    width = max(e - b for b, e in ranges)
    mine = torch.zeros(width, dtype=local.dtype, device=local.device)
    mine[:len(local)] = local
    gathered = torch.empty((world, width), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(gathered.view(-1), mine, group=group)
    else:
        dist.all_gather([gathered[k] for k in range(world)], mine, group=group)
    out = torch.empty(total, dtype=local.dtype, device=local.device)
    for k, (b, e) in enumerate(ranges):
        out[b:e] = gathered[k, :e - b]
    assert ranges[rank][1] - ranges[rank][0] == len(local)
    return out


class PendingGather:
    """An all-gather of per-rank contiguous pieces in flight (NCCL: on the communicator's stream, beside whatever
    the caller enqueues next on the library's stream).  finish() / finish_into() wait for it and place the
    pieces."""

    def __init__(self, local, ranges, group=None):
        import torch
        import torch.distributed as dist
        self.ranges, self.local, self.handle, self.gathered = ranges, local, None, None
        world = len(ranges)
        if world == 1 or not (dist.is_available() and dist.is_initialized()):
            assert len(local) == ranges[0][1] - ranges[0][0]
            return
        rank = dist.get_rank(group)
        assert ranges[rank][1] - ranges[rank][0] == len(local)
        width = max(e - b for b, e in ranges)
        mine = torch.zeros(width, dtype=local.dtype, device=local.device)
        mine[:len(local)] = local
        self.gathered = torch.empty((world, width), dtype=local.dtype, device=local.device)
        if dist.get_backend(group) == "nccl":
            self.handle = dist.all_gather_into_tensor(self.gathered.view(-1), mine, group=group, async_op=True)
        else:
            self.handle = dist.all_gather([self.gathered[k] for k in range(world)], mine, group=group, async_op=True)
        self._mine = mine   # keep alive until the collective is done

    def finish_into(self, out):
        if self.gathered is None:
            b, e = self.ranges[0]
            out[b:e] = self.local
            return out
        self.handle.wait()
        for k, (b, e) in enumerate(self.ranges):
            out[b:e] = self.gathered[k, :e - b]
        return out

    def finish(self, total):
        import torch
        return self.finish_into(torch.empty(total, dtype=self.local.dtype, device=self.local.device))


class StreamWorker:
    """One shard of an exactly sharded stream on this process's GPU (torch CUDA tensors, C-ABI calls).  A rank
    that owns several shards (bounded memory per step) shares pipeline / Harvest objects between them:
    `share` = a worker whose pipeline this one uses too (their stages then run one after the other), `harvests_of`
    = a worker whose Harvest objects it uses.  `stream`: a torch stream every stage of this worker is enqueued on
    WITHOUT waiting for it (the caller orders streams and waits once; workers of different pipelines then overlap
    on the GPU); None = the library's own stream, every stage waited for (simple, serial)."""

    def __init__(self, plan, shard, harvest_option=None, cheaptrick_option=None, d4c_option=None, share=None,
                 harvests_of=None, stream=None):
        import torch
        import worldb200 as wb
        self.wb, self.torch, self.plan, self.rank, self.stream = wb, torch, plan, shard, stream
        self.harvest_option = harvest_option if harvest_option is not None else wb.HarvestOption()
        assert abs(self.harvest_option.frame_period - plan.fp) < 1e-12
        if share is not None:
            self.pipe = share.pipe
        else:
            self.pipe = wb.Pipeline(plan.fs, self.harvest_option, cheaptrick_option, d4c_option)
            self.pipe.set_fresh_rng(True)                       # the stream is one reference process
        src = harvests_of if harvests_of is not None else share
        if src is not None:
            self.harvests, self.streams = src.harvests, src.streams
        else:
            # several Harvest objects on their own streams: consecutive segments overlap (the one-CTA contour tail
            # of one runs beside the band-pass bank of the other)
            self.harvests = [wb.Harvest(plan.fs, self.harvest_option) for _ in range(HARVESTS_IN_FLIGHT)]
            self.streams = [torch.cuda.Stream() for _ in range(HARVESTS_IN_FLIGHT)]
        self.harvest = self.harvests[0]
        assert self.pipe.fft_size == plan.fft_size, "plan was made for another FFT size"
        self.bins = plan.fft_size // 2 + 1

    # every stage: (serial mode) wait for torch's stream, run on the library's, wait for the device;
    # (stream mode) allocate and run on self.stream, wait for nothing
    def _enter(self):
        if self.stream is None:
            self.torch.cuda.current_stream().synchronize()   # inputs made by torch ops; the library has its own stream
            return None
        return ctypes.c_void_p(self.stream.cuda_stream)

    def _leave(self):
        if self.stream is None:
            self.wb.device_synchronize()

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def harvest_local(self, d_x):
        """f0 of the frames this shard owns (whole-stream frame grid)."""
        return harvest_shards([self], d_x)

    def begin(self, d_f0_all, samples=None):
        """whole-stream bookkeeping (frame times, randn() seed, time base + pulse list): once per pipeline.
        `samples`: the sample range the shards of this pipeline will synthesise (default: this shard's) -- the
        pulse list is built for it only."""
        st = self._enter()
        self.d_f0_all = d_f0_all
        sa, sb = samples if samples is not None else self.plan.samples[self.rank]
        self.wb._check(self.wb.lib().wb_pipeline_stream_begin_range_dev(self.pipe._h, d_f0_all.data_ptr(), self.plan.f0_length,
                                                                        self.plan.out_length, int(sa), int(sb), st),
                       "wb_pipeline_stream_begin_range_dev")

    def envelope(self, d_x, d_ap0_all):
        """CheapTrick + Love Train for the rows this shard needs; writes its entries of d_ap0_all."""
        st = self._enter()
        torch = self.torch
        ra, rb = self.plan.rows[self.rank]
        with self._on_stream():
            self.d_sp = torch.empty((rb - ra, self.bins), dtype=torch.float64, device=d_x.device)
        self.wb._check(self.wb.lib().wb_pipeline_stream_envelope_dev(self.pipe._h, d_x.data_ptr(), self.plan.n, self.d_f0_all.data_ptr(),
                                                                     self.plan.f0_length, ra, rb, self.d_sp.data_ptr(),
                                                                     d_ap0_all.data_ptr(), st), "wb_pipeline_stream_envelope_dev")
        self._leave()

    def lovetrain(self, d_x, d_ap0_all):
        """first half of envelope(): Love Train for the rows this shard needs; writes its entries of d_ap0_all"""
        st = self._enter()
        ra, rb = self.plan.rows[self.rank]
        self.wb._check(self.wb.lib().wb_pipeline_stream_lovetrain_dev(self.pipe._h, d_x.data_ptr(), self.plan.n, self.d_f0_all.data_ptr(),
                                                                      self.plan.f0_length, ra, rb, d_ap0_all.data_ptr(), st),
                       "wb_pipeline_stream_lovetrain_dev")
        self._leave()

    def cheaptrick(self, d_x, sync=True):
        """second half of envelope(): the spectral envelope rows this shard needs"""
        st = self._enter()
        torch = self.torch
        ra, rb = self.plan.rows[self.rank]
        with self._on_stream():
            self.d_sp = torch.empty((rb - ra, self.bins), dtype=torch.float64, device=d_x.device)
        self.wb._check(self.wb.lib().wb_pipeline_stream_cheaptrick_dev(self.pipe._h, d_x.data_ptr(), self.plan.n, self.d_f0_all.data_ptr(),
                                                                       self.plan.f0_length, ra, rb, self.d_sp.data_ptr(), st),
                       "wb_pipeline_stream_cheaptrick_dev")
        if sync:
            self._leave()

    def aperiodicity(self, d_x, d_ap0_all):
        st = self._enter()
        torch = self.torch
        ra, rb = self.plan.rows[self.rank]
        with self._on_stream():
            self.d_ap = torch.empty((rb - ra, self.bins), dtype=torch.float64, device=d_x.device)
        self.wb._check(self.wb.lib().wb_pipeline_stream_aperiodicity_dev(self.pipe._h, d_x.data_ptr(), self.plan.n, self.d_f0_all.data_ptr(),
                                                                         d_ap0_all.data_ptr(), self.plan.f0_length, ra, rb,
                                                                         self.d_ap.data_ptr(), st), "wb_pipeline_stream_aperiodicity_dev")
        self._leave()

    def synthesis(self):
        st = self._enter()
        torch = self.torch
        ra, rb = self.plan.rows[self.rank]
        sa, sb = self.plan.samples[self.rank]
        with self._on_stream():
            d_y = torch.zeros(max(1, sb - sa), dtype=torch.float64, device=self.d_sp.device)
        self.wb._check(self.wb.lib().wb_pipeline_stream_synthesis_dev(self.pipe._h, self.plan.f0_length, self.d_sp.data_ptr(),
                                                                      self.d_ap.data_ptr(), ra, rb - ra, self.plan.out_length, sa, sb,
                                                                      d_y.data_ptr(), st), "wb_pipeline_stream_synthesis_dev")
        self._leave()
        return d_y[:sb - sa]

    def end(self):
        st = self._enter()
        self.wb._check(self.wb.lib().wb_pipeline_stream_end_dev(self.pipe._h, st), "wb_pipeline_stream_end_dev")
        if self.stream is None:
            self.wb.device_synchronize()
            self.pipe.check_errors()     # (conditions a kernel of the asynchronous calls above could not handle)
        else:
            self.pipe.check_errors(self.stream.cuda_stream)    # (waits for the stream)

    def owned_rows(self, d_rows):
        """the rows of d_sp / d_ap this shard owns (without the halo rows it computed for its pulses)"""
        ra, _ = self.plan.rows[self.rank]
        fb, fe = self.plan.frames[self.rank]
        return d_rows[fb - ra:fe - ra]


def harvest_shards(workers, d_x):
    """f0 of the frames the given shards of one rank own (consecutive shards sharing their Harvest objects), as one
    tensor.  The padded segments of all of them go round-robin over HARVESTS_IN_FLIGHT Harvest objects / streams
    and are waited for once."""
    w0 = workers[0]
    torch, wb, plan = w0.torch, w0.wb, w0.plan
    torch.cuda.current_stream().synchronize()   # inputs made by torch ops; the library calls run on their own streams
    L = wb.lib()
    fb, fe = plan.frames[workers[0].rank][0], plan.frames[workers[-1].rank][1]
    out = torch.zeros(fe - fb, dtype=torch.float64, device=d_x.device)
    done = []
    k = len(w0.harvests)
    for seg in [s for w in workers for s in plan.segments[w.rank]]:
        pa, pb = seg["padded"]
        n = pb - pa
        h, st = w0.harvests[len(done) % k], w0.streams[len(done) % k]
        n_local = h.getSamples(plan.fs, n)
        d_t = torch.empty(n_local, dtype=torch.float64, device=d_x.device)
        d_f = torch.empty(n_local, dtype=torch.float64, device=d_x.device)
        wb._check(L.wb_harvest_compute_dev(h._h, d_x.data_ptr() + 8 * pa, n, d_t.data_ptr(), d_f.data_ptr(),
                                           ctypes.c_void_p(st.cuda_stream)), "wb_harvest_compute_dev")
        done.append((seg, d_f, d_t))
    for h, st in zip(w0.harvests, w0.streams):
        h.check_errors(st.cuda_stream)       # (waits for the stream)
    for seg, d_f, _ in done:
        cfb, cfe = seg["frames"]
        off = seg["frame_offset"]
        out[cfb - fb:cfe - fb] = d_f[cfb - off:cfe - off]
    return out


def process_stream_exact(d_x, fs, harvest_option=None, cheaptrick_option=None, d4c_option=None, segment_seconds=120,
                         halo_seconds=2, group=None, d_f0_all=None, shards_per_rank=1, keep_rows=True, timings=None,
                         state=None, pipelines_in_flight=1):
    """Analysis + re-synthesis of one long stream (a float64 CUDA tensor every rank holds) sharded over the
    ranks of `group`.  Returns dict(f0 [whole], y [whole], sp, ap [this rank's rows], frames, plan).
    Pass d_f0_all to skip Harvest (e.g. a contour computed elsewhere).  shards_per_rank > 1 processes the
    rank's share in that many pieces (bounds the scratch memory of one step; same results).  keep_rows=False
    drops each shard's sp / ap rows once its samples are synthesised.  `timings`: a dict that receives CUDA-event
    milliseconds per phase.  `state`: a dict the caller keeps between calls on streams of the same shape; the
    pipeline objects (and with them their device workspaces: several GB for long streams) are then reused
    instead of being allocated and freed by every call.  pipelines_in_flight: how many groups of the rank's shards
    run concurrently (one pipeline object and CUDA stream each; at most shards_per_rank).  Measured on one hour of
    48 kHz audio on one B200: 1 / 2 / 4 / 8 groups take 464.5 / 465.9 / 467.4 / 473.4 ms -- the frame kernels of a
    long shard each fill the register file of every SM, so a second pipeline finds nothing to overlap with (unlike
    batches of short utterances, cf. BatchPipeline); the default is therefore 1."""
    import torch
    import torch.distributed as dist
    import worldb200 as wb
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    k = max(1, int(shards_per_rank))
    hopt = harvest_option if harvest_option is not None else wb.HarvestOption()
    copt = cheaptrick_option if cheaptrick_option is not None else wb.CheapTrickOption()
    fft_size = copt.fft_size if copt.fft_size else wb.CheapTrick.getFFTSizeForCheapTrick(fs, copt.f0_floor)
    plan = StreamPlan(d_x.numel(), fs, world * k, hopt.frame_period, fft_size, segment_seconds, halo_seconds, hopt.target_fs)
    mine = list(range(rank * k, (rank + 1) * k))
    # The rank's shards form `n_groups` contiguous groups; a group shares one pipeline object (its shards run one
    # after the other: bounded scratch) on its own CUDA stream; groups may overlap on the GPU (see the docstring).
    n_groups = max(1, min(k, int(pipelines_in_flight)))
    key = (plan.n, plan.fs, world, k, n_groups, plan.fp, plan.fft_size, int(segment_seconds), int(halo_seconds))
    if state is not None and state.get("key") == key:
        groups = state["groups"]
    else:
        groups = []
        for g in range(n_groups):
            members = mine[(k * g) // n_groups:(k * (g + 1)) // n_groups]
            stream = torch.cuda.Stream()
            ws = []
            for s in members:
                ws.append(StreamWorker(plan, s, hopt, cheaptrick_option, d4c_option, share=ws[0] if ws else None,
                                       harvests_of=groups[0][0] if groups else (ws[0] if ws else None), stream=stream))
            groups.append(ws)
        if state is not None:
            state.update(key=key, groups=groups)
    workers = [w for ws in groups for w in ws]
    main = torch.cuda.current_stream()

    def fork():      # the group streams see what torch's stream has produced so far
        for ws in groups:
            ws[0].stream.wait_stream(main)

    def join():      # ... and torch's stream what the groups have produced
        for ws in groups:
            main.wait_stream(ws[0].stream)

    def round_robin():   # shard order that keeps every group's stream fed: first shards of all groups, then the second ...
        for j in range(max(len(ws) for ws in groups)):
            for ws in groups:
                if j < len(ws):
                    yield ws[j]

    # per-rank contiguous ranges for the exchanges
    rank_frames = [(plan.frames[r * k][0], plan.frames[(r + 1) * k - 1][1]) for r in range(world)]
    marks = []

    def mark(name):
        if timings is not None:
            ev = torch.cuda.Event(enable_timing=True)
            wb.device_synchronize()
            ev.record()
            marks.append((name, ev))

    mark("start")
    external_f0 = d_f0_all is not None
    if d_f0_all is None:
        local_f0 = harvest_shards(workers, d_x)
        mark("harvest")
        d_f0_all = gather_ranges(local_f0, rank_frames, plan.f0_length, group)                   # exchange 1
        mark("gather_f0")
    # (an external contour is the caller's: size the pulse buffers by its maximum instead of Harvest's ceiling)
    f0_bound = float(d_f0_all.max().item()) + 1.0 if external_f0 else 0.0
    d_ap0 = torch.zeros(plan.f0_length, dtype=torch.float64, device=d_x.device)
    fork()
    for ws in groups:
        ws[0].pipe.set_stream_f0_bound(f0_bound)
        ws[0].begin(d_f0_all, (plan.samples[ws[0].rank][0], plan.samples[ws[-1].rank][1]))
        for w in ws[1:]:
            w.d_f0_all = d_f0_all
    # Love Train first: its decisions (8 bytes per frame) travel while CheapTrick -- the longer half of the
    # envelope work -- runs, so the ranks do not sit in the exchange waiting for the slowest of them
    for w in round_robin():
        w.lovetrain(d_x, d_ap0)
    join()
    mark("lovetrain")
    fb, fe = rank_frames[rank]
    pending_ap0 = PendingGather(d_ap0[fb:fe].clone(), rank_frames, group)                        # exchange 2 (in flight)
    for w in round_robin():
        w.cheaptrick(d_x)
    join()
    mark("cheaptrick")
    d_ap0 = pending_ap0.finish(plan.f0_length)
    main.synchronize()
    mark("wait_ap0")
    fork()
    # shard by shard: aperiodicity + synthesis, and the shard's samples go into the exchange (the stitch) while
    # the other shards compute; only the last exchanges are exposed
    done = {}
    pending, sps, aps = [], [], []

    def exchange_finished_shards():      # in shard order (the same on every rank), as far as they have been enqueued
        while len(pending) < k and mine[len(pending)] in done:
            j = len(pending)
            w = workers[j]
            y, ev = done.pop(mine[j])
            main.wait_event(ev)
            pending.append(PendingGather(y, [plan.samples[r * k + j] for r in range(world)], group))   # exchange 3
            if keep_rows:
                sps.append(w.owned_rows(w.d_sp))
                aps.append(w.owned_rows(w.d_ap))
            w.d_sp = w.d_ap = None

    for w in round_robin():
        w.aperiodicity(d_x, d_ap0)
        y = w.synthesis()
        ev = torch.cuda.Event()
        ev.record(w.stream)
        done[w.rank] = (y, ev)
        exchange_finished_shards()
    for ws in groups:
        ws[0].end()
    mark("aperiodicity_synthesis")
    d_y = torch.empty(plan.out_length, dtype=torch.float64, device=d_x.device)
    for pg in pending:
        pg.finish_into(d_y)
    main.synchronize()
    mark("stitch_tail")
    if timings is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            timings[name] = e0.elapsed_time(e1)
        timings["total"] = marks[0][1].elapsed_time(marks[-1][1])
    return {"f0": d_f0_all, "y": d_y, "sp": torch.cat(sps) if sps else None, "ap": torch.cat(aps) if aps else None,
            "frames": (fb, fe), "plan": plan}


def simulate_stream_ranks(d_x, fs, world, harvest_option=None, cheaptrick_option=None, d4c_option=None, segment_seconds=120,
                          halo_seconds=2, d_f0_all=None):
    """The same computation as process_stream_exact() with `world` VIRTUAL ranks run one after the other in
    this process (one worker object each, as separate processes would have); the exchanges become plain
    concatenations.  Used by the single-GPU tests."""
    import torch
    import worldb200 as wb
    hopt = harvest_option if harvest_option is not None else wb.HarvestOption()
    copt = cheaptrick_option if cheaptrick_option is not None else wb.CheapTrickOption()
    fft_size = copt.fft_size if copt.fft_size else wb.CheapTrick.getFFTSizeForCheapTrick(fs, copt.f0_floor)
    plan = StreamPlan(d_x.numel(), fs, world, hopt.frame_period, fft_size, segment_seconds, halo_seconds, hopt.target_fs)
    workers = [StreamWorker(plan, k, hopt, cheaptrick_option, d4c_option) for k in range(world)]
    if d_f0_all is None:
        d_f0_all = torch.cat([w.harvest_local(d_x) for w in workers])
    assert d_f0_all.numel() == plan.f0_length
    d_ap0 = torch.zeros(plan.f0_length, dtype=torch.float64, device=d_x.device)
    pieces = []
    for w in workers:
        w.begin(d_f0_all)
        mine = torch.zeros_like(d_ap0)
        w.envelope(d_x, mine)
        fb, fe = plan.frames[w.rank]
        pieces.append(mine[fb:fe])
    d_ap0 = torch.cat(pieces)
    for w in workers:
        w.aperiodicity(d_x, d_ap0)
    d_y = torch.cat([w.synthesis() for w in workers])
    return {"f0": d_f0_all, "y": d_y, "sp": torch.cat([w.owned_rows(w.d_sp) for w in workers]),
            "ap": torch.cat([w.owned_rows(w.d_ap) for w in workers]), "ap0": d_ap0, "plan": plan, "workers": workers}
