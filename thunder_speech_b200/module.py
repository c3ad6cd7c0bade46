"""``BaseCTCModule`` forward / predict (mirrors ``src/thunder/module.py:25-100``) without Lightning.

    module = CTCModule(encoder, decoder, audio_transform, text_transform)
    logits, out_lengths = module(audio, lengths)        # forward, module.py:74-86
    texts = module.predict(audio)                       # module.py:88-100

``forward`` keeps activations in the kernels' bf16 row layout from the feature kernel to the decoder GEMM and
returns fp32 logits ``[B, V, T']`` like the reference.  ``predict`` adds the greedy CTC kernel and detokenises
on the host after a single device-to-host copy.  ``training_step`` / ``validation_step`` / ``configure_optimizers``
(module.py:102-192) keep the reference's signatures on top of ``thunder_speech_b200.train`` (kernel backward behind one
autograd node); Lightning itself (logging, trainer hooks) is not part of this package.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops

__all__ = ["CTCModule", "BaseCTCModule"]

#: default number of sub-batch chains per captured inference graph (see CTCModule.graph_chains_for)
DEFAULT_CHAINS = 1


class CTCModule(nn.Module):
    MAX_IN_PLACE_GRAPHS = 8     # captured graphs kept for caller-owned input buffers (oldest evicted first)

    def __init__(self, encoder: nn.Module, decoder: nn.Module, audio_transform: nn.Module,
                 text_transform: nn.Module, optimizer_class=torch.optim.AdamW, optimizer_kwargs: Optional[Dict] = None,
                 lr_scheduler_class=None, lr_scheduler_kwargs: Optional[Dict] = None,
                 encoder_final_dimension: Optional[int] = None):
        """Same constructor as ``BaseCTCModule`` (src/thunder/module.py:26-64), argument for argument."""
        super().__init__()
        self.audio_transform = audio_transform
        self.encoder = encoder
        self.decoder = decoder
        self.text_transform = text_transform
        self.optimizer_class = optimizer_class
        self.optimizer_kwargs = dict(optimizer_kwargs or {})
        self.lr_scheduler_class = lr_scheduler_class
        self.lr_scheduler_kwargs = dict(lr_scheduler_kwargs or {})
        self.lr_scheduler_interval = self.lr_scheduler_kwargs.pop("interval", "step")
        self.encoder_final_dimension = encoder_final_dimension
        # metrics of validation_step (module.py:67-68 keeps torchmetrics' CharErrorRate / WordErrorRate)
        from .metrics import CharErrorRate, WordErrorRate

        self.validation_cer = CharErrorRate()
        self.validation_wer = WordErrorRate()
        self.precision: Optional[str] = None    # None = the package default (thunder_speech_b200.get_default_precision())
        self._dec_cache = None
        self._graphs: Dict[tuple, "_PredictGraph"] = {}
        self._pipes: Dict[tuple, "_StreamPipe"] = {}
        self._graph_weights = None      # flat list of the tensors the captured graphs bake in (built lazily)
        self._graph_fingerprint = None
        #: independent sub-batch CHAINS inside one captured graph (None = chosen from the batch size, see graph_chains_for)
        self.graph_chains: Optional[int] = None

    # -- row format of the inference path ---------------------------------------------------------
    def set_precision(self, precision: Optional[str]) -> "CTCModule":
        """``"bf16"`` / ``"fp16"`` activation rows and GEMM operands for forward / predict (None = package default); see
        ``thunder_speech_b200.set_default_precision``.  Captured graphs are per precision."""
        from . import row_dtype

        row_dtype(precision)     # validates
        self.precision = precision
        return self

    def _row_dtype(self) -> torch.dtype:
        from . import row_dtype

        return row_dtype(self.precision)

    # -- captured graphs are only valid for the weights they were captured with ----------------
    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name in ("encoder", "decoder", "audio_transform") and "_graphs" in self.__dict__:
            self.invalidate_graphs()

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda() / .half(): parameters move, graphs die
        out = super()._apply(fn, *args, **kwargs)
        if "_graphs" in self.__dict__:
            self.invalidate_graphs()
        return out

    def invalidate_graphs(self) -> None:
        """Drop every captured inference graph and serving pipeline (they are rebuilt on the next graphed call).  Called
        automatically when a weight / buffer of the encoder, decoder or front-end changes in place (``load_state_dict``,
        an optimizer step, ``CTCTrainStep``), moves (``.to()``), or a sub-module is replaced; call it yourself after
        replacing an individual ``nn.Parameter`` object deep inside a block."""
        self._graphs.clear()
        self._pipes.clear()
        self._graph_weights = None
        self._graph_fingerprint = None

    def _weights_fingerprint(self) -> tuple:
        """``(data_ptr, _version)`` of every parameter and buffer a captured forward reads (folded plans and the decoder
        operands are derived from them at capture time, so the graph bakes in tensors that die when these change)."""
        if self._graph_weights is None:
            mods = [self.encoder, self.decoder, self.audio_transform]
            self._graph_weights = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
        return tuple((t.data_ptr(), t._version) for t in self._graph_weights)

    def _check_graphs_current(self) -> None:
        fp = self._weights_fingerprint()
        if fp != self._graph_fingerprint:
            if self._graph_fingerprint is not None:
                self._graphs.clear()
                self._pipes.clear()
            self._graph_fingerprint = fp

    # -- decoder parameters as GEMM operands ----------------------------------------------------
    def _decoder_operands(self) -> Tuple[Tensor, Tensor]:
        dec = self.decoder
        if not isinstance(dec, nn.Conv1d) or dec.kernel_size[0] != 1:
            raise NotImplementedError("only conv1d_decoder (1x1 Conv1d with bias) is implemented")
        dt = self._row_dtype()
        key = (dec.weight.data_ptr(), dec.weight._version, dec.bias.data_ptr(), dec.bias._version, dt)
        if self._dec_cache is None or self._dec_cache[0] != key:
            w = dec.weight.detach()[:, :, 0].float()
            if dt == torch.float16:
                w = w.clamp(-65504.0, 65504.0)
            w = w.to(dt).contiguous()
            b = dec.bias.detach().float().contiguous()
            self._dec_cache = (key, w, b)
        return self._dec_cache[1], self._dec_cache[2]

    def _logits_rows(self, x: Tensor, lengths: Tensor):
        """audio -> (fp32 logits [B,V,T'], T', device i32 lengths, i64 feature lengths)."""
        N = x.shape[-1]
        hop = self.audio_transform[1].hop_length
        F = 1 + N // hop
        feats, feat_len = self.audio_transform.features(x, lengths, bf16_pitch=ops.row_pitch(F),
                                                        f16=self._row_dtype() == torch.float16)
        l32 = feat_len.to(torch.int32)
        rows, T, l32 = self.encoder.forward_rows(feats, F, l32)
        w, b = self._decoder_operands()
        logits = ops.pw_gemm(w, rows, None, None, T, b, None, True, False, None, None, None)
        return logits, T, l32, feat_len

    def forward(self, x: Tensor, lengths: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
        """``(audio[B,N], lengths[B]) -> (logits[B,V,T'] f32, out_lengths[B] i64)`` (module.py:74-86)."""
        if self.training:
            raise NotImplementedError("forward() is the inference path: call .eval(), or use training_step() / "
                                      "thunder_speech_b200.train.CTCTrainStep for the train()-mode step")
        with torch.no_grad():
            logits, T, l32, _ = self._logits_rows(x, lengths)
            return logits, l32.to(torch.int64)

    @torch.no_grad()
    def predict_ids(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """Device part of ``predict``: ``(argmax ids [B,T'], collapsed ids [B,T'] (-1 padded), counts [B])``."""
        lengths = torch.full((x.shape[0],), x.shape[-1], device=x.device, dtype=torch.int64)  # module.py:98
        logits, T, _, _ = self._logits_rows(x, lengths)
        return ops.ctc_greedy(logits, T, -1)

    @torch.no_grad()
    def predict(self, x: Tensor) -> List[str]:
        """Audio ``[batch, time]`` -> transcriptions (module.py:88-100): every frame is decoded, lengths = N."""
        _, col, cnt = self.predict_ids(x)
        return self.text_transform.decode_collapsed(col, cnt)

    # -- CUDA-graph replay of predict_ids for a fixed (batch, samples) shape -----------------------
    @torch.no_grad()
    def predict_ids_graphed(self, x: Tensor, in_place: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
        """Same result as :meth:`predict_ids`; the ~200 kernel launches of one forward are captured once per
        input shape into a CUDA graph and replayed (launch-bound otherwise at small batch).  By default the audio is
        copied into the graph's own input buffer; ``in_place=True`` declares ``x`` a persistent device buffer owned by the
        caller (a DMA staging buffer): the graph captured for it reads ``x`` where it lies, one graph per such buffer."""
        self._check_graphs_current()    # stale weights must never be replayed (train / load_state_dict in between)
        key = (x.shape[0], x.shape[1], x.data_ptr() if in_place else None, self._row_dtype(), self.graph_chains_for(x.shape[0]))
        g = self._graphs.get(key)
        if g is None:
            if in_place:    # a caller that passes a fresh tensor every time must not accumulate graphs (and their memory pools)
                stale = [k for k in self._graphs if k[2] is not None]
                for k in stale[:max(0, len(stale) - (self.MAX_IN_PLACE_GRAPHS - 1))]:
                    del self._graphs[k]
            g = _PredictGraph(self, x, in_place)
            self._graphs[key] = g
        return g.replay(x)

    def graph_chains_for(self, batch: int) -> int:
        """How many independent utterance chains a captured forward of ``batch`` utterances is split into.  Utterances
        never interact in eval mode, so the graph may run the network over disjoint slices of the batch on parallel
        branches: while one chain's kernel drains (last tiles, TMA store flush) or sets up (TMEM, barriers, first loads),
        the other chain's kernel has the SMs -- the per-kernel fixed cost that dominates small batches is overlapped
        instead of serialised.  ``graph_chains`` (attribute) or THUNDER_B200_CHAINS overrides the choice."""
        import os

        n = self.graph_chains
        if n is None and os.environ.get("THUNDER_B200_CHAINS"):
            n = int(os.environ["THUNDER_B200_CHAINS"])
        if n is None:
            n = DEFAULT_CHAINS
        return max(1, min(int(n), batch))

    @torch.no_grad()
    def _predict_ids_chained(self, x: Tensor, chains: int) -> Tuple[Tensor, Tensor, Tensor]:
        """``predict_ids`` over ``chains`` contiguous slices of the batch, each on its own stream (forked from / joined to
        the current one, so it can be captured into one graph); results concatenated in batch order."""
        if chains <= 1:
            return self.predict_ids(x)
        B = x.shape[0]
        cur = torch.cuda.current_stream()
        pool = self.__dict__.setdefault("_chain_streams", [])
        while len(pool) < chains - 1:
            pool.append(torch.cuda.Stream(device=x.device))
        side = pool[:chains - 1]
        for s in side:      # fork BEFORE chain 0 queues anything on the current stream
            s.wait_stream(cur)
        parts = []
        for i in range(chains):
            lo, hi = B * i // chains, B * (i + 1) // chains
            with torch.cuda.stream(cur if i == 0 else side[i - 1]):
                parts.append(self.predict_ids(x[lo:hi]))
        for s in side:
            cur.wait_stream(s)
        return tuple(torch.cat(p, 0) for p in zip(*parts))

    def training_step(self, batch, batch_idx: int = 0) -> Tensor:
        """``BaseCTCModule.training_step`` (src/thunder/module.py:102-127): ``batch = (audio, audio_lengths, texts)`` -> mean
        CTC loss.  The returned scalar is connected to torch's autograd graph: ``loss.backward()`` (what Lightning's
        automatic optimisation calls next) delivers the kernel-computed gradients to ``param.grad`` of every encoder /
        decoder parameter, so any torch optimizer / LR scheduler / gradient-clipping hook keeps working.  The explicit,
        faster loop (single-launch AdamW, no autograd node) is ``thunder_speech_b200.train.CTCTrainStep.step``."""
        from .train import CTCTrainStep

        step = self.__dict__.get("_b200_step")
        if step is None:
            step = self.__dict__["_b200_step"] = CTCTrainStep(self)
        audio, audio_lengths, texts = batch
        y, y_lengths = self.text_transform.encode(texts, device=audio.device)
        return step.autograd_loss(audio, audio_lengths, y, y_lengths)

    def _update_special_optimizer_arg(self, original_kwargs: Dict) -> Dict:
        """``total_steps_arg="<name>"`` is replaced by ``<name>=trainer.estimated_stepping_batches`` (module.py:165-171)."""
        updated = dict(original_kwargs)
        total_steps_arg = updated.pop("total_steps_arg", None)
        if total_steps_arg:
            trainer = getattr(self, "trainer", None)
            if trainer is None:
                raise RuntimeError("total_steps_arg needs a Lightning trainer attached to the module (self.trainer)")
            updated[total_steps_arg] = trainer.estimated_stepping_batches
        return updated

    def configure_optimizers(self):
        """``BaseCTCModule.configure_optimizers`` (module.py:173-192): ``optimizer_class`` over the trainable parameters,
        optionally with ``lr_scheduler_class`` in Lightning's dictionary form.  The optimizer consumes ``param.grad`` as
        delivered by ``training_step(...).backward()``."""
        optimizer = self.optimizer_class(filter(lambda p: p.requires_grad, self.parameters()),
                                         **self._update_special_optimizer_arg(self.optimizer_kwargs))
        if not self.lr_scheduler_class:
            return optimizer
        scheduler = self.lr_scheduler_class(optimizer, **self._update_special_optimizer_arg(self.lr_scheduler_kwargs))
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": self.lr_scheduler_interval}}

    @torch.no_grad()
    def validation_step(self, batch, batch_idx: int = 0) -> Tensor:
        """``BaseCTCModule.validation_step`` (src/thunder/module.py:130-163), eval() mode: ``batch = (audio, audio_lengths,
        texts)`` -> mean CTC loss (the CTC kernel on the fp32 logits with the encoder's output lengths, like
        ``calculate_ctc``); greedy transcriptions of EVERY frame (``probabilities.argmax(1)``) and the label strings update
        ``validation_cer`` / ``validation_wer``."""
        from .train import ctc_loss

        audio, audio_lengths, texts = batch
        y, y_lengths = self.text_transform.encode(texts, device=audio.device)
        if self.training:
            raise NotImplementedError("validation_step() runs the eval() forward (Lightning switches to eval for validation)")
        logits, T, l32, _ = self._logits_rows(audio, audio_lengths)
        loss_b, _ = ctc_loss(logits, T, l32, y, y_lengths.to(torch.int64), self.text_transform.vocab.blank_idx)
        _, col, cnt = ops.ctc_greedy(logits, T, -1)
        decoded_preds = self.text_transform.decode_collapsed(col, cnt)
        decoded_targets = self.text_transform.decode_prediction(y, remove_repeated=False)
        self.validation_cer(decoded_preds, decoded_targets)
        self.validation_wer(decoded_preds, decoded_targets)
        return loss_b.mean()

    @torch.no_grad()
    def predict_graphed(self, x: Tensor) -> List[str]:
        _, col, cnt = self.predict_ids_graphed(x)
        return self.text_transform.decode_collapsed(col, cnt)

    @torch.no_grad()
    def to_torchscript(self, example_audio: Tensor, example_lengths: Optional[Tensor] = None,
                       file_path: Optional[str] = None) -> "torch.jit.ScriptModule":
        """TorchScript export with the reference's two entry points (Lightning's ``to_torchscript`` on ``BaseCTCModule``:
        ``forward`` plus the ``@torch.jit.export``-ed ``predict``, src/thunder/module.py:74-100):

            ts = module.to_torchscript(example_audio)
            logits, out_lengths = ts(audio, lengths)
            texts = ts.predict(audio)                  # List[str], greedy CTC incl. detokenisation, inside TorchScript

        The kernels enter the graph by TRACING (a sequence of ``torch.ops.thunder_b200.*`` calls; any batch size, but the frame
        counts are Python ints at trace time, so the export is specialised to the audio LENGTH of the example); the
        detokeniser is a scripted module with the reference's string rules.  ``torch.jit.script`` of the whole model is not
        offered -- the reference's own front-end does not script on torch 2.x either.  Loading in a fresh
        process needs ``import thunder_speech_b200.ops`` first so that the custom ops are registered."""
        import warnings

        if example_lengths is None:
            example_lengths = torch.full((example_audio.shape[0],), example_audio.shape[-1], device=example_audio.device)
        was_training = self.training
        self.eval()
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            traced = torch.jit.trace_module(self, {"forward": (example_audio, example_lengths),
                                                   "_predict_collapsed": (example_audio,)}, check_trace=False)
            v = self.text_transform.vocab
            specials = [v.blank_token, v.pad_token] + [t for t in (v.start_token, v.end_token) if t is not None]
            exported = torch.jit.script(_TorchScriptExport(traced, list(v.itos), specials))
        self.train(was_training)
        if file_path is not None:
            torch.jit.save(exported, file_path)
        return exported

    def _predict_collapsed(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """Traceable device part of ``predict``: (collapsed ids ``[B, T']`` padded with -1, counts ``[B]``)."""
        _, col, cnt = self.predict_ids(x)
        return col, cnt

    def predict_stream(self, batches: Iterable[Tensor], depth: int = 3, remove_dc: bool = False) -> Iterator[List[str]]:
        """Serving loop over HOST batches ``[B, N]`` of one fixed shape (pinned memory for true overlap): float32 audio
        like ``predict``, or **int16 PCM** as it comes out of a wav file -- half the host-to-device bytes; the samples are
        scaled by 1/32768 on the device like ``torchaudio.load`` does on the host (``ts_pcm_ingest``; ``remove_dc`` also
        subtracts the per-utterance mean like ``AudioFileLoader.preprocess_audio``, src/thunder/data/dataset.py:50-77).  The
        host-to-device copies run on a copy stream ahead of the compute (CUDA-graph replay) and the detokenisation of
        finished batches runs on the CPU meanwhile.  ``depth`` batches are in flight: with 3 the copy of batch i+1 is
        queued before the host waits for batch i-1, so it has a whole step to finish even when the PCIe transfer takes
        almost as long as the compute (with 2 it only starts after batch i-1 was detokenised).  Yields one
        ``List[str]`` per batch, in order."""
        depth = max(2, int(depth))
        pipe = None
        pending = []
        for i, xb in enumerate(batches):
            if pipe is None:
                # staging buffers, pinned result buffers and the captured graph are kept per (shape, depth): building
                # them costs ~20 ms (cudaHostAlloc, graph warm-up), which a short stream would pay on every call
                key = (tuple(xb.shape), depth, xb.dtype, bool(remove_dc))
                pipe = self._pipes.get(key)
                if pipe is None:
                    pipe = self._pipes[key] = _StreamPipe(self, xb, depth, remove_dc)
                pipe.reset()
            pending.append(pipe.submit(i, xb))
            if len(pending) == depth:
                yield pipe.collect(pending.pop(0))
        while pending:
            yield pipe.collect(pending.pop(0))


class _TorchScriptExport(nn.Module):
    """What ``CTCModule.to_torchscript`` returns (scripted): ``forward`` = the traced kernel sequence, ``predict`` = traced
    greedy-CTC ids + the reference's detokenisation rules (text_processing/transform.py:93-122, vocab.py:85-130)."""

    def __init__(self, traced: torch.nn.Module, itos: List[str], specials: List[str]):
        super().__init__()
        self.traced = traced
        self.itos = itos
        self.specials = specials

    def forward(self, x: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
        return self.traced(x, lengths)

    @torch.jit.export
    def predict(self, x: Tensor) -> List[str]:
        col, cnt = self.traced._predict_collapsed(x)
        col = col.cpu()
        cnt = cnt.cpu()
        out: List[str] = []
        for b in range(col.size(0)):
            s = ""
            for t in range(int(cnt[b])):
                s += self.itos[int(col[b, t])]
            s = s.replace("\u2581", " ")      # sentencepiece word boundary
            s = s.replace("|", " ")           # huggingface word boundary
            for sp in self.specials:          # blank, pad (, start, end) removed as SUBSTRINGS, in the reference's order
                s = s.replace(sp, "")
            out.append(s)
        return out


class _StreamPipe:
    """Multi-buffered H2D / compute / D2H pipeline behind :meth:`CTCModule.predict_stream`."""

    def __init__(self, module: CTCModule, example: Tensor, depth: int = 3, remove_dc: bool = False):
        self.m = module
        self.depth = depth
        self.remove_dc = remove_dc
        if example.dtype not in (torch.float32, torch.int16):
            raise TypeError("predict_stream: batches must be float32 audio or int16 PCM")
        self.pcm = example.dtype == torch.int16
        dev = next(module.encoder.parameters()).device
        B, N = example.shape
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty((B, N), device=dev, dtype=torch.float32) for _ in range(depth)]
        # int16 PCM lands here (half the PCIe bytes) and one ingest kernel converts it into the graph's float32 input
        self.stage16 = [torch.empty((B, 1, N), device=dev, dtype=torch.int16) for _ in range(depth)] if self.pcm else None
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.stage_free = [torch.cuda.Event() for _ in range(depth)]
        self.d2h_done = [torch.cuda.Event() for _ in range(depth)]
        # one in-place graph per staging buffer (no D2D copy) while that fits the module's graph budget
        self.in_place = depth <= module.MAX_IN_PLACE_GRAPHS - 2
        for st in self.stage:
            _, col, cnt = module.predict_ids_graphed(st, in_place=self.in_place)
        torch.cuda.synchronize(dev)
        self.host_col = [torch.empty(col.shape, dtype=col.dtype).pin_memory() for _ in range(depth)]
        self.host_cnt = [torch.empty(cnt.shape, dtype=cnt.dtype).pin_memory() for _ in range(depth)]

    def reset(self) -> None:
        """Start of a new stream over the same buffers: nothing of the previous one may still be in flight."""
        torch.cuda.current_stream().synchronize()
        self.copy_stream.synchronize()

    def submit(self, i: int, xb: Tensor) -> int:
        s = i % self.depth
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if i >= self.depth:
                self.copy_stream.wait_event(self.stage_free[s])
            if self.pcm:
                self.stage16[s].copy_(xb.view(self.stage16[s].shape), non_blocking=True)
            else:
                self.stage[s].copy_(xb, non_blocking=True)
            self.h2d_done[s].record(self.copy_stream)
        cur.wait_event(self.h2d_done[s])
        if self.pcm:
            from .data import pcm_ingest

            pcm_ingest(self.stage16[s], None, False, self.remove_dc, out=self.stage[s])
        _, col, cnt = self.m.predict_ids_graphed(self.stage[s], in_place=self.in_place)
        self.stage_free[s].record(cur)
        self.host_col[s].copy_(col, non_blocking=True)
        self.host_cnt[s].copy_(cnt, non_blocking=True)
        self.d2h_done[s].record(cur)
        return s

    def collect(self, s: int) -> List[str]:
        self.d2h_done[s].synchronize()
        return self.m.text_transform.decode_collapsed(self.host_col[s], self.host_cnt[s])


class _PredictGraph:
    def __init__(self, module: CTCModule, example: Tensor, in_place: bool = False):
        from . import _lib

        self.static_in = example if in_place else example.clone()
        self.chains = module.graph_chains_for(example.shape[0])
        # warm-up on a side stream (sets kernel attributes, builds plans), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                module._predict_ids_chained(self.static_in, self.chains)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = module._predict_ids_chained(self.static_in, self.chains)
        self.kernels_per_replay = _lib.launch_count() - n0
        self.replays = 0

    def replay(self, x: Tensor):
        if x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_out


BaseCTCModule = CTCModule
