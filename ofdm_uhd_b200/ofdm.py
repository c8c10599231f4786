"""ofdm_mod / ofdm_demod with the constructor options, ``send_pkt`` and rx-callback semantics of the
reference's ofdm.py (/root/reference/ofdm.py:38-305), running on B200 kernels instead of a GNU Radio
flowgraph.

The reference classes are ``gr.hier_block2`` objects with one complex stream port; a scheduler thread per
block moves the samples.  Here there is no flowgraph: packets queued by ``send_pkt`` are modulated as one
batch by ``flush()`` (also triggered when the queue is full and by ``eof``) and pushed to whatever was
``connect()``-ed; ``ofdm_demod.feed(samples)`` runs the receiver on a buffer and delivers
``callback(ok, payload)`` from a daemon watcher thread, in arrival order -- as ofdm.py:290-305 does.
"""
from __future__ import annotations

import math
import queue
import threading
import traceback

import numpy as np

try:
    from . import ofdm_packet_utils
    from .engine import OfdmEngine, MODS
except ImportError:                      # flat import, like the reference's script directory
    import ofdm_packet_utils
    from engine import OfdmEngine, MODS


# ------------------------------------------------------------------------------------------------
# the sliver of gr.message / gr.msg_queue the callers touch (ofdm.py:139-148)
# ------------------------------------------------------------------------------------------------
class message:
    def __init__(self, type=0, arg1=0.0, arg2=0.0, data=b""):
        self._type, self._arg1, self._arg2, self._data = type, arg1, arg2, bytes(data)

    def type(self):
        return self._type

    def arg1(self):
        return self._arg1

    def arg2(self):
        return self._arg2

    def length(self):
        return len(self._data)

    def to_string(self):
        return self._data


def message_from_string(s, type=0, arg1=0.0, arg2=0.0):
    return message(type, arg1, arg2, s)


class msg_queue:
    """Thread-safe FIFO with gr.msg_queue's method names; ``limit`` 0 means unbounded."""

    def __init__(self, limit=0, on_full=None):
        self._q = queue.Queue()
        self._limit = limit
        self._on_full = on_full

    def limit(self):
        return self._limit

    def count(self):
        return self._q.qsize()

    def empty_p(self):
        return self._q.empty()

    def full_p(self):
        return self._limit != 0 and self._q.qsize() >= self._limit

    def insert_tail(self, msg):
        # gr.msg_queue blocks the producer here until the mapper has drained a slot; the drain is ours to run
        if self.full_p() and self._on_full is not None:
            self._on_full()
        self._q.put(msg)

    def delete_head(self):
        return self._q.get()

    def delete_head_nowait(self):
        try:
            return self._q.get_nowait()
        except queue.Empty:
            return None

    def flush(self):
        while self.delete_head_nowait() is not None:
            pass


class _pkt_input:
    """Stands where digital.ofdm_mapper_bcv stood: callers only reach ``_pkt_input.msgq()`` (ofdm.py:148)."""

    def __init__(self, msgq_limit, on_full):
        self._msgq = msg_queue(msgq_limit, on_full)

    def msgq(self):
        return self._msgq


def _known_preamble(fft_length, occupied_tones):
    """ofdm.py:70-87: the known symbol with odd absolute bins zeroed, and its padded vector."""
    zeros_on_left = int(math.ceil((fft_length - occupied_tones) / 2.0))
    ksfreq = list(known_symbols_4512_3[0:occupied_tones])
    for i in range(len(ksfreq)):
        if (zeros_on_left + i) & 1:
            ksfreq[i] = 0
    padded = fft_length * [0, ]
    padded[zeros_on_left: zeros_on_left + occupied_tones] = ksfreq
    return ksfreq, padded


# /////////////////////////////////////////////////////////////////////////////
#                   mod/demod with packets as i/o
# /////////////////////////////////////////////////////////////////////////////

class ofdm_mod:
    """
    Modulates an OFDM stream. Based on the options fft_length, occupied_tones, and cp_length, this
    creates OFDM symbols using a specified modulation option.  Send packets by calling send_pkt.
    """

    def __init__(self, options, msgq_limit=2, pad_for_usrp=True, batch_limit=4096, device=None, pad_seed=0,
                 carrier_map=None):
        """
        @param options: pass modulation options from higher layers (fft length, occupied tones, etc.)
        @param msgq_limit: maximum number of messages in message queue (kept for compatibility: the queue is
               drained by ``flush()``, which runs automatically every ``batch_limit`` packets and on eof)
        @param pad_for_usrp: If true, packets are padded such that they end up a multiple of 128 samples
        """
        self._pad_for_usrp = pad_for_usrp
        self._modulation = options.modulation
        self._fft_length = options.fft_length
        self._occupied_tones = options.occupied_tones
        self._cp_length = options.cp_length
        self._msgq_limit = msgq_limit
        self._batch_limit = max(1, int(batch_limit))

        ksfreq, padded = _known_preamble(self._fft_length, self._occupied_tones)
        self.preambles = (padded,)
        arity = MODS[self._modulation]                      # KeyError for an unknown modulation (ofdm.py:92)
        self._arity = arity
        # ofdm_mod scales by 1/sqrt(N) only; transmit_path owns the amplitude stage (transmit_path.py:48)
        self._engine = OfdmEngine(self._fft_length, self._occupied_tones, self._cp_length, self._modulation,
                                  tx_amplitude=1.0, device=device, pad_seed=pad_seed, carrier_map=carrier_map)
        self._device, self._pad_seed = device, pad_seed
        self._pkt_input = _pkt_input(self._batch_limit, self.flush)
        self._sinks = []
        self._frames_sent = 0
        self._eof = False
        self._lock = threading.Lock()
        # attribute names of the reference blocks (ofdm.py:111-114): the stages are fused into one kernel
        self.ifft = self.cp_adder = self.scale = self._engine
        if options.verbose:
            self._print_verbage()
        self._log = bool(getattr(options, "log", False))

    # -- flowgraph replacement ------------------------------------------------------------------
    def reset_carrier_map(self, carrier_map):
        """What the reference's custom mapper offered as ``_pkt_input.reset_carrier_map`` (the call commented
        out at transmit_path.py:67): pending packets go out with the old map, later ones use the new one."""
        self.flush()
        amp = getattr(self, "_amp", 1.0)
        old = self._engine
        self._engine = OfdmEngine(self._fft_length, self._occupied_tones, self._cp_length, self._modulation,
                                  tx_amplitude=amp, device=self._device, pad_seed=self._pad_seed,
                                  carrier_map=carrier_map)
        self.ifft = self.cp_adder = self.scale = self._engine
        old.close()

    def connect(self, sink):
        """sink: callable(samples) or object with .feed(samples); samples is a complex64 cuda tensor."""
        self._sinks.append(sink)
        return sink

    def modulate(self):
        """Drain the packet queue: returns a complex64 cuda tensor (frames back to back), or None."""
        import torch
        with self._lock:
            pkts = []
            q = self._pkt_input.msgq()
            while True:
                m = q.delete_head_nowait()
                if m is None:
                    break
                if m.type() == 1:
                    self._eof = True
                    continue
                pkts.append(m.to_string())
            if not pkts:
                return None
            off = np.zeros(len(pkts) + 1, dtype=np.int64)
            np.cumsum([len(p) for p in pkts], out=off[1:])
            host = torch.frombuffer(bytearray(b"".join(pkts)), dtype=torch.uint8)
            dev = host.to(self._engine.dev, non_blocking=False)
            taps = {} if self._log else None
            out = self._engine.modulate(dev, off, first_frame=self._frames_sent, taps=taps)
            self._frames_sent += len(pkts)
            if self._log:
                # the reference's four file sinks (ofdm.py:123-131), raw interleaved float32 I/Q; truncated by the first
                # batch, appended to afterwards.  ofdm_cp_adder_c.dat holds what leaves this block (the reference taps it
                # in front of its 1/sqrt(N) scale stage)
                mode = "ab" if getattr(self, "_log_started", False) else "wb"
                self._log_started = True
                for name, t in (("ofdm_mapper_c.dat", taps["mapper"]), ("ofdm_preambles.dat", taps["preambles"]),
                                ("ofdm_ifft_c.dat", taps["ifft"]), ("ofdm_cp_adder_c.dat", out)):
                    with open(name, mode) as fh:
                        t.cpu().numpy().tofile(fh)
            return out

    def flush(self):
        out = self.modulate()
        if out is not None:
            for s in self._sinks:
                (s.feed if hasattr(s, "feed") else s)(out)
        return out

    def send_pkt(self, payload='', eof=False):
        """
        Send the payload.

        @param payload: data to send
        @type payload: bytes (str is encoded latin-1)
        """
        if eof:
            msg = message(1)          # tell self._pkt_input we're not sending any more packets
        else:
            pkt = ofdm_packet_utils.make_packet(payload, 1, 1, self._pad_for_usrp, whitening=True)
            msg = message_from_string(pkt)
        self._pkt_input.msgq().insert_tail(msg)
        if eof:
            self.flush()

    def send_pkts(self, payloads):
        """Bulk twin of :meth:`send_pkt`: the whole list is framed on the device (make_packets_kernel: header, CRC-32,
        whitening -- what ofdm_packet_utils.make_packet does per packet on the host) and modulated in one pass; packets
        queued by earlier send_pkt calls go out first.  Raises ValueError where make_packet would."""
        import torch
        self.flush()
        payloads = [p.encode("latin-1") if isinstance(p, str) else bytes(p) for p in payloads]
        if not payloads:
            return None
        with self._lock:
            off = np.zeros(len(payloads) + 1, dtype=np.int64)
            np.cumsum([len(p) for p in payloads], out=off[1:])
            plan = self._engine.tx_plan(off, pad_for_usrp=self._pad_for_usrp)
            blob = b"".join(payloads)
            host = torch.frombuffer(bytearray(blob), dtype=torch.uint8) if blob else torch.zeros(1, dtype=torch.uint8)
            out = self._engine.tx_run(plan, host.to(self._engine.dev), first_frame=self._frames_sent)
            self._frames_sent += len(payloads)
        for s in self._sinks:
            (s.feed if hasattr(s, "feed") else s)(out)
        return out

    @staticmethod
    def add_options(normal, expert):
        """
        Adds OFDM-specific options to the Options Parser
        """
        normal.add_option("-m", "--modulation", type="string", default="bpsk",
                          help="set modulation type (bpsk, qpsk, 8psk, qam{16,64}) [default=%default]")
        expert.add_option("", "--fft-length", type="int", default=512,
                          help="set the number of FFT bins [default=%default]")
        expert.add_option("", "--occupied-tones", type="int", default=200,
                          help="set the number of occupied FFT bins [default=%default]")
        expert.add_option("", "--cp-length", type="int", default=128,
                          help="set the number of bits in the cyclic prefix [default=%default]")

    def _print_verbage(self):
        """
        Prints information about the OFDM modulator
        """
        print("\nOFDM Modulator:")
        print("Modulation Type: %s" % (self._modulation))
        print("FFT length:      %3d" % (self._fft_length))
        print("Occupied Tones:  %3d" % (self._occupied_tones))
        print("CP length:       %3d" % (self._cp_length))


class ofdm_demod:
    """
    Demodulates a received OFDM stream. Based on the options fft_length, occupied_tones, and cp_length,
    this performs synchronization, FFT, and demodulation of incoming OFDM symbols and passes packets up
    to a higher layer via the callback.
    """

    def __init__(self, options, callback=None, device=None, max_pkt_bytes=4096, carrier_map=None):
        """
        @param options: pass modulation options from higher layers (fft length, occupied tones, etc.)
        @param callback:  function of two args: ok, payload
        @type callback: ok: bool; payload: bytes
        """
        self._rcvd_pktq = msg_queue()          # holds packets from the PHY
        self._modulation = options.modulation
        self._fft_length = options.fft_length
        self._occupied_tones = options.occupied_tones
        self._cp_length = options.cp_length
        self._snr = options.snr                # parsed, unused by the live "pn" synchroniser (ofdm.py:208)

        ksfreq, _ = _known_preamble(self._fft_length, self._occupied_tones)
        self.preambles = (ksfreq,)
        self._arity = MODS[self._modulation]
        self._engine = OfdmEngine(self._fft_length, self._occupied_tones, self._cp_length, self._modulation,
                                  device=device, max_pkt_bytes=max_pkt_bytes, carrier_map=carrier_map)
        self.ofdm_recv = self.ofdm_demod = self._engine
        self._log = bool(getattr(options, "log", False))
        # SYNC of ofdm_receiver.py~:89-119 is a source-level switch there ("pn" live, "fixed" for testing only);
        # here it is read from the options object, with the reference's hard-coded fixed-mode parameters as defaults
        self._sync = getattr(options, "sync", "pn")
        self._sync_nsymbols = int(getattr(options, "sync_nsymbols", 18))
        self._sync_freq_offset = float(getattr(options, "sync_freq_offset", 0.0))
        if self._sync not in ("pn", "ml", "pnac", "fixed"):
            raise ValueError("sync %r: ofdm_receiver.py names 'pn', 'ml', 'pnac' and 'fixed'" % (self._sync,))
        if options.verbose:
            self._print_verbage()
        self._watcher = _queue_watcher_thread(self._rcvd_pktq, callback)
        self._batch_callback = None
        self.last = None
        # continuous-stream mode (feed_stream): tail of the stream seen so far, its absolute position, and the
        # absolute start of the last frame handed to the callback
        self._carry = None
        self._carry_abs = 0
        self._stream_job = None                       # dense pass queued on the device, not yet handed to the callback
        self._stream_sets = ({}, {})                  # its two alternating buffer sets
        self._stream_passes = 0
        self._pass_streams = None                     # two CUDA streams the dense passes alternate on
        self._dense_rate = None                       # (messages, payload bytes) per sample seen so far: sizes the D2H
        self._last_abs_start = None
        # feed_stream batches small buffers: the receiver runs once this many new samples are pending (a receive pass is
        # ~17 kernel launches and re-reads the carried tail, so per-buffer passes on radio-sized buffers would be
        # launch-bound); 0 = run on every call
        self.stream_batch_samples = int(getattr(options, "stream_batch_samples", 0))
        self._stream_pending = []
        self._stream_pending_n = 0

    def stream_carry_samples(self):
        """Samples of the past that feed_stream() keeps in front of every new buffer: the detector's IIR warm-up
        (24 576 samples), the channel-filter delay, and the longest frame the header can announce (4095 + 9 bytes),
        so that a frame cut by a buffer boundary is demodulated whole by the next call."""
        eng = self._engine
        longest = 1 + -(-8 * (4095 + 9) // (eng.ncar * eng.nbits))
        return 24576 + eng.ntaps + (longest + 1) * eng.L

    def feed_stream(self, samples, max_frames=None, flush=False):
        """feed() for a source delivered in consecutive buffers (a file read in pieces, a radio): the receiver runs
        on [carried tail | new samples] and only frames that start after the last delivered one reach the callback,
        so every frame is delivered once, in order, wherever the buffer boundaries fall.  (The reference's flowgraph
        is continuous; feed() by itself treats each buffer as a separate stream.)

        With ``stream_batch_samples`` > 0 the buffers are queued (on the device) until that many new samples are
        pending -- or ``flush`` / flush_stream() -- and then go through the receiver in one pass: the cost of a pass
        (launches + the carried tail of stream_carry_samples()) is amortised over the batch, so a source that delivers
        64 k-sample buffers runs at the whole-stream rate; delivery is delayed by at most one batch.  With a batch
        callback set (set_batch_callback) a pass is handed over while the NEXT one runs on the device -- one more batch
        of delay, no host wait; flush_stream() delivers everything outstanding."""
        torch = self._engine.torch
        if samples is not None:
            # (this prologue is all a call costs until a batch is full: keep it to a few attribute reads)
            if type(samples) is not torch.Tensor:
                samples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.complex64))
            if not samples.is_cuda:
                samples = samples.to(self._engine.dev)
            n_new = samples.numel()
            if n_new:
                self._stream_pending.append(samples if samples.is_contiguous() else samples.contiguous())
                n_new += self._stream_pending_n
                self._stream_pending_n = n_new
                if not flush and n_new < self.stream_batch_samples:
                    return None
        if not flush and self._stream_pending_n < self.stream_batch_samples:
            return None
        dense = self._batch_callback is not None and self._sync == "pn" and not self._log
        if not self._stream_pending:
            return self._stream_finish(True) if dense else None
        parts = ([self._carry] if self._carry is not None else []) + self._stream_pending
        abs0 = self._carry_abs
        L = self._engine.L
        if dense:
            return self._stream_pass_dense(parts, abs0, max_frames, flush)
        buf = parts[0] if len(parts) == 1 else torch.cat(parts)
        self._stream_pending, self._stream_pending_n = [], 0
        res = self.feed(buf, max_frames=max_frames, _deliver=False)
        # a frame still open at the end of the buffer (the sink ran out of vectors) is left to the next call
        delivered = []
        for k, f in enumerate(res.msg_frames):
            start = abs0 + int(res.frame_start[f])
            if self._last_abs_start is not None and start <= self._last_abs_start + L:
                continue
            self._last_abs_start = start
            delivered.append(res.packets[k])
        for ok, payload in delivered:
            self._rcvd_pktq.insert_tail(_rx_message(ok, payload))
        keep = min(self.stream_carry_samples(), buf.numel())
        self._carry = buf[buf.numel() - keep:].clone()
        self._carry_abs = abs0 + buf.numel() - keep
        res.stream_delivered = delivered
        return res

    def _stream_pass_dense(self, parts, abs0, max_frames, flush):
        """One pass of feed_stream with the dense hand-over, one pass behind: this pass is queued on the device
        (receiver, packing, device->host copies sized by what earlier passes brought back) and the PREVIOUS pass, long
        finished, is handed to the callback, so the host never waits for the GPU while the source keeps delivering
        buffers.  Consecutive passes alternate between two buffer sets AND two CUDA streams: a pass depends on its
        predecessor only through the carried sample tail (copied out on the caller's stream before the receiver
        starts), so the launch-bound end of pass k (trigger compaction, plan, liveness, CRC, packing, copies) runs
        under the stream kernels of pass k + 1."""
        eng = self._engine
        torch = eng.torch
        k = self._stream_passes
        self._stream_passes += 1
        if self._pass_streams is None:
            self._pass_streams = (torch.cuda.Stream(device=eng.dev), torch.cuda.Stream(device=eng.dev))
        ps = self._pass_streams[k & 1]
        # [carried tail | queued buffers] and the next tail are put together on the CALLER's stream (where the buffers
        # were produced, and where the next pass will look for the tail); the pass stream picks the batch up from there
        buf = parts[0] if len(parts) == 1 else torch.cat(parts)
        self._stream_pending, self._stream_pending_n = [], 0
        n = int(buf.numel())
        keep = min(self.stream_carry_samples(), n)
        self._carry = buf[n - keep:].clone()
        self._carry_abs = abs0 + n - keep
        ps.wait_stream(torch.cuda.current_stream(eng.dev))
        buf.record_stream(ps)
        with torch.cuda.stream(ps):
            job = {"k": k, "abs0": abs0, "buf": buf, "n": n}
            slot = self._stream_sets[k & 1]
            cap = max(n, slot.get("cap", 0))
            L = eng.L
            mf = max_frames if max_frames is not None else max(64, cap // L + 64)
            if slot.get("bufs") is None or slot["cap"] < cap or slot["mf"] != mf:
                slot.update(bufs=eng.rx_alloc(cap, max_frames=mf, fresh=True), cap=cap, mf=mf)
            bufs = eng.demodulate_async(buf, slot["bufs"])
            em, eb = self._dense_expect(n)
            job["bufs"] = bufs
            job["ticket"] = eng.deliver_begin(bufs, expect_msgs=em, expect_bytes=eb, frame_starts=True)
        r = self._stream_finish(False)                           # the pass before this one
        self._stream_job = job
        if flush:
            r = self._stream_finish(True)
        return r if r is not None else {"stream_delivered": 0, "n_msgs": 0, "pending_pass": job["k"]}

    def _dense_expect(self, n):
        """Sizes of the dense hand-over's device->host copies for a pass over n samples: what earlier passes brought
        back per sample, + 25 % (None, None = everything, before anything is known).  A pass that exceeds them costs
        one more, exactly sized copy (deliver_end(complete=True)), never a lost message."""
        if self._dense_rate is None:
            return None, None
        return (int(1.25 * self._dense_rate[0] * n) + 64,
                int(1.25 * self._dense_rate[1] * n) + 64 * self._engine.pkt_stride)

    def _dense_learn(self, r, n):
        m = r["n_msgs"]
        if m and n > 0:
            old = self._dense_rate or (0.0, 0.0)
            self._dense_rate = (max(old[0], m / n), max(old[1], float(r["off"][m]) / n))

    def _stream_finish(self, _unused=None):
        """Hand the queued dense pass (if any) to the batch callback: its messages come back as one byte array +
        offsets; the ones already delivered by the pass before (frames inside the carried tail) are a prefix."""
        job, self._stream_job = self._stream_job, None
        if job is None:
            return None
        eng = self._engine
        r = eng.deliver_end(job["ticket"], complete=True)
        n = r["n_msgs"]
        L = eng.L
        self._dense_learn(r, job["n"])
        starts = job["abs0"] + r["frame_start"] if n else np.zeros(0, np.int64)
        k0 = 0
        if self._last_abs_start is not None and n:
            k0 = int(np.searchsorted(starts, self._last_abs_start + L, side="right"))
        if n > k0:
            self._last_abs_start = int(starts[-1])
            off = r["off"]
            self._batch_callback(r["ok"][k0:], r["data"][int(off[k0]):int(off[n])], off[k0:] - off[k0])
        r["stream_delivered"] = n - k0
        self.last = r
        return r

    def reset_stream(self):
        """Forget the continuous stream seen so far (a new capture starts): anything still queued is dropped."""
        self._carry, self._carry_abs, self._last_abs_start = None, 0, None
        self._stream_pending, self._stream_pending_n, self._stream_job = [], 0, None
        self._stream_passes = 0                       # the next capture starts on the first buffer set again

    def flush_stream(self, max_frames=None):
        """Run the receiver on whatever feed_stream() has queued."""
        return self.feed_stream(None, max_frames=max_frames, flush=True)

    def set_batch_callback(self, fn):
        """Bulk twin of the per-packet callback: with ``fn`` set, every feed() hands ALL the packets it produced to
        ``fn(ok, data, offsets)`` in one call -- ok: bool array [n], data: uint8 array of the payload || crc bytes back
        to back, offsets: int64 [n + 1] (payload of packet k = data[offsets[k] : offsets[k+1] - 4]) -- straight from the
        device's dense hand-over (ofdm_rx_compact) instead of one Python message per packet.  ``None`` restores the
        per-packet path."""
        self._batch_callback = fn

    def feed(self, samples, max_frames=None, _deliver=True):
        """Run the receiver on one buffer of complex64 samples (cuda tensor, or host array copied to the
        device) and queue every packet the frame sink produced for the watcher thread.  Each call is a
        self-contained stream (filter history, detector average and NCO phase start from zero)."""
        import torch
        if not isinstance(samples, torch.Tensor):
            samples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.complex64))
        if samples.device.type != "cuda":
            samples = samples.to(self._engine.dev)
        samples = samples.contiguous()
        if self._batch_callback is not None and _deliver and self._sync == "pn" and not self._log:
            bufs = self._engine.demodulate_async(samples, max_frames=max_frames)
            em, eb = self._dense_expect(int(samples.numel()))
            r = self._engine.deliver_end(self._engine.deliver_begin(bufs, expect_msgs=em, expect_bytes=eb), complete=True)
            n = r["n_msgs"]
            self._dense_learn(r, int(samples.numel()))
            self._batch_callback(r["ok"], r["data"][:int(r["off"][n])] if n else r["data"][:0], r["off"])
            self.last = r
            return r
        if self._sync == "fixed":
            res = self._engine.demodulate_fixed(samples, self._sync_nsymbols, self._sync_freq_offset, max_frames=max_frames)
        elif self._sync in ("ml", "pnac"):
            res = self._engine.collect(self._engine.demodulate_async(samples, sync=self._sync, snr_db=float(self._snr),
                                                                     max_frames=max_frames))
        elif self._log:
            res = self._feed_logged(samples, max_frames)
        else:
            res = self._engine.demodulate(samples, max_frames=max_frames)
        self.last = res
        if _deliver:
            for ok, payload in res.packets:
                self._rcvd_pktq.insert_tail(_rx_message(ok, payload))
        return res

    def _feed_logged(self, samples, max_frames):
        """options.log: dump the stage taps with the reference's file names and raw layout (ofdm.py:253-254,
        ofdm_receiver.py~:144-152; interleaved float32 I/Q, utils/read_complex_binary.m:39-46): chan_filt, fft_out,
        frame_acq, found_corr, sampler, sigmix, nco and ofdm_frame_sink.  Files are truncated by the first feed() and
        appended to afterwards."""
        eng = self._engine
        n = int(samples.numel())
        nvec = n // eng.L + 64
        bufs = eng.rx_alloc(n, max_frames=max_frames, taps="all", max_vectors=nvec)
        for k in ("eq_syms", "sym_idx", "derot_syms", "fft_out", "sampler_out"):
            bufs[k].zero_()
        res = eng.collect(eng.demodulate_async(samples, bufs))
        nco, sigmix = eng.nco_taps(bufs, n)
        mode = "ab" if getattr(self, "_log_started", False) else "wb"
        self._log_started = True
        flags = np.concatenate([np.concatenate([[1], np.zeros(int(j), np.uint8)]) for j in res.frame_ndata]).astype(np.uint8) \
            if res.n_frames else np.zeros(0, np.uint8)
        nv = len(flags)
        derot = bufs["derot_syms"][:nv * eng.ncar].cpu().numpy().reshape(nv, eng.ncar)
        wide = np.zeros((nv, eng.occ), dtype=np.complex64)
        wide[:, :eng.ncar] = derot
        for name, arr in (("ofdm_receiver-chan_filt_c.dat", eng.ws_view(bufs, 0, n).cpu().numpy()),
                          ("ofdm_receiver-fft_out_c.dat", bufs["fft_out"][:nv * eng.N].cpu().numpy()),
                          ("ofdm_receiver-frame_acq_c.dat", bufs["eq_syms"][:nv * eng.occ].cpu().numpy()),
                          ("ofdm_receiver-found_corr_b.dat", flags),
                          ("ofdm_receiver-sampler_c.dat", bufs["sampler_out"][:nv * eng.N].cpu().numpy()),
                          ("ofdm_receiver-sigmix_c.dat", sigmix.cpu().numpy()),
                          ("ofdm_receiver-nco_c.dat", nco.cpu().numpy()),
                          ("ofdm_frame_sink_c.dat", wide[np.abs(derot).sum(axis=1) > 0])):   # only the vectors the sink demapped
            with open(name, mode) as f:
                arr.tofile(f)
        return res

    def wait(self, timeout=None):
        """Block until the watcher has delivered everything queued so far."""
        self._watcher.drain(timeout)

    @staticmethod
    def add_options(normal, expert):
        """
        Adds OFDM-specific options to the Options Parser
        """
        normal.add_option("-m", "--modulation", type="string", default="bpsk",
                          help="set modulation type (bpsk or qpsk) [default=%default]")
        expert.add_option("", "--fft-length", type="int", default=512,
                          help="set the number of FFT bins [default=%default]")
        expert.add_option("", "--occupied-tones", type="int", default=200,
                          help="set the number of occupied FFT bins [default=%default]")
        expert.add_option("", "--cp-length", type="int", default=128,
                          help="set the number of bits in the cyclic prefix [default=%default]")

    def _print_verbage(self):
        """
        Prints information about the OFDM demodulator
        """
        print("\nOFDM Demodulator:")
        print("Modulation Type: %s" % (self._modulation))
        print("FFT length:      %3d" % (self._fft_length))
        print("Occupied Tones:  %3d" % (self._occupied_tones))
        print("CP length:       %3d" % (self._cp_length))


class _rx_message(message):
    """A received packet whose dewhitening and CRC check already ran on the device (ofdm_rx_finish)."""

    def __init__(self, ok, payload):
        message.__init__(self, 0, 0, 0, payload)
        self.ok = bool(ok)


class _queue_watcher_thread(threading.Thread):
    """ofdm.py:290-305: pops received packets and fires callback(ok, payload), one at a time, in order,
    for bad CRCs too."""

    def __init__(self, rcvd_pktq, callback):
        threading.Thread.__init__(self)
        self.daemon = True
        self.rcvd_pktq = rcvd_pktq
        self.callback = callback
        self.keep_running = True
        self._pending = 0
        self.errors = 0
        self._cv = threading.Condition()
        _insert = rcvd_pktq.insert_tail

        def counted_insert(msg):
            with self._cv:
                self._pending += 1
            _insert(msg)
        rcvd_pktq.insert_tail = counted_insert
        self.start()

    def run(self):
        while self.keep_running:
            msg = self.rcvd_pktq.delete_head()
            try:
                if isinstance(msg, _rx_message):
                    ok, payload = msg.ok, msg.to_string()
                else:
                    ok, payload = ofdm_packet_utils.unmake_packet(msg.to_string())
                if self.callback:
                    self.callback(ok, payload)
            except Exception:
                # a failing user callback must not end the thread: later packets would never be delivered and
                # wait() would block for ever on _pending
                self.errors += 1
                traceback.print_exc()
            finally:
                with self._cv:
                    self._pending -= 1
                    self._cv.notify_all()

    def drain(self, timeout=None):
        with self._cv:
            self._cv.wait_for(lambda: self._pending == 0, timeout)


# The reference's 4512 known symbols (ofdm.py:310-325, "i = [2*random.randint(0,1)-1 for i in range(4512)]"),
# kept as a packed bit string (bit = 1 <=> +1, LSB first); tests/test_tables.py pins the SHA-256 of the
# expanded list against the reference file.
def _expand_known_symbols():
    import base64
    packed = np.frombuffer(base64.b85decode(
        "G%Yn+z+C((y^1&>9SHr$Q%^4)KPP8_N!hz@_mT_hMPvTxbqaQkWSi~Wu8B!QR)zkQ)?zPm4aH;<^;58S?h8z>gq=eL_VMpm?DBrxa2xPsU>54"
        ";i5Ru+gfA;GH$bMdh_xJhbN6kQ9)ryE|IZ?)lhgjCkOlAekgb&5m$FbNK%Pmv|0=(<v(p0LR%@q-)c0D`wL6(X^8?*2YAJeZE+~CFpi;G2eEiD"
        "Odn=*7&a>x+PX~&MIuu~7%bh%cQomBX^95JLBM0i)MRP!ug~mrlXB7BpNLLXlee6tz#FW&H?aNNkYkly63%RC$RBee%Aat=1+U^cPDrIKt|8XH"
        "S3gA?LIbFBq63Ljc%e$T{6tVZn#a|2pXi-U}YxJByY%$CtU{u*{3~>*h>h31-=c!WIGcg<p3~_<>J_8m%OCM-bl@>7_oL&`GTCGH;-X2I3?Yl"
        "}b)m2yTpd5)>DM{#Xip|l$!o3)kCK`q9U<LEsg-M(yP~uH(k4z5oFj20cdCeIh^m;Gcvohp>V6iA65@;>N%zpw7P<7^|Fg@F7nw$CRx~H&uR!"
        "4k~i5XL-hcE<7jm>uHxDL<<8tR3xCLADjp691EO9&P=s~@I7i>0L!lW5T9Zhj22RarGG8+Ti-M`_e;Xci|P-7!lnOKc(gdR8LNP`45$hx8#Hay"
        "1_Nc3F)TUXg|=B}+BSL*;>0dPw{g&!t;!18_z1Te0emdM_YY"), dtype=np.uint8)
    bits = np.unpackbits(packed, bitorder="little")[:4512]
    return [int(2 * int(b) - 1) for b in bits]


known_symbols_4512_3 = _expand_known_symbols()
