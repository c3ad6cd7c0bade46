"""CPU tests of everything that runs on the host: C-ABI library load + exported symbols, module shells and
their state_dict contract, BatchNorm folding, mel/window tables, greedy-decode string rules, length formulas,
batch sharding over gloo (world_size 2).  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_numpy as R
from thunder_speech_b200 import _lib, synth
from thunder_speech_b200.blocks import conv1d_decoder, conv_out_length, get_same_padding, lengths_to_mask
from thunder_speech_b200.citrinet.blocks import CitrinetBlock, CitrinetEncoder
from thunder_speech_b200.fused import build_block_plan
from thunder_speech_b200.parallel import shard_bounds
from thunder_speech_b200.quartznet.blocks import QuartznetBlock, QuartznetEncoder
from thunder_speech_b200.quartznet.transform import FilterbankFeatures, mel_filterbank
from thunder_speech_b200.text_processing import BatchTextTransformer, Vocabulary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "thunder_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", src)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()                       # raises loudly if the .so is missing
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/thunder_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == declared
    assert b"sm_100a" in lib.ts_version()
    assert _lib.row_pitch(751) == 768 and _lib.row_pitch(64) == 64 and _lib.row_pitch(1) == 64
    assert _lib.launch_count() >= 0
    # argument errors are reported without touching a GPU
    rc = lib.ts_set_option(b"no_such_option", 1)
    assert rc == _lib.TS_ERR_INVALID and b"unknown option" in lib.ts_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.ts_dw_conv(1, 1, 4, 50, 64, 1, 3, 2, 2, 1, None, 0, 1, 64, None), "ts_dw_conv")  # stride & dilation > 1


def test_library_has_no_libcuda_link_dependency():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out          # driver entry points are resolved at run time


def test_ops_refuse_cpu_tensors():
    from thunder_speech_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pack_rows(torch.zeros(1, 2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FilterbankFeatures().eval()(torch.zeros(1, 2000), torch.tensor([2000]))


def test_filterbank_tables_and_state_dict_contract(golden_features):
    fb = FilterbankFeatures(nfilt=64)
    assert list(fb.state_dict().keys()) == ["1.window", "2.layer.0.fb"]
    assert fb[2].layer[0].fb.shape == (1, 64, 257)
    assert np.abs(fb[2].layer[0].fb[0].numpy() - golden_features["fb64"]).max() < 5e-6 * golden_features["fb64"].max()
    assert np.abs(mel_filterbank(257, 80, 16000, 0.0, 8000.0).numpy() - golden_features["fb80"]).max() < 2e-7
    assert np.abs(fb[1].window.numpy() - golden_features["window320"]).max() < 1e-7
    t = fb._device_tables(torch.device("cpu"))
    assert (t["win_lo"], t["win_hi"]) == (97, 415)                            # hann endpoints are exactly zero
    # sparse bank reproduces the dense one
    dense = np.zeros((64, 257), np.float32)
    for m in range(64):
        s, c, o = int(t["mel_start"][m]), int(t["mel_count"][m]), int(t["mel_off"][m])
        dense[m, s:s + c] = t["mel_w"][o:o + c].numpy()
    assert np.array_equal(dense, fb[2].layer[0].fb[0].numpy())
    assert torch.equal(fb[1].get_sequence_length(torch.tensor([0, 159, 160, 320000])), torch.tensor([1, 1, 2, 2001]))
    with pytest.raises(ValueError):
        FilterbankFeatures(n_window_stride=0)
    with pytest.raises(ValueError):
        FilterbankFeatures(num_cutout_masks=2, num_freq_masks=1)


def test_length_helpers_match_reference(golden_helpers):
    g = golden_helpers
    assert np.array_equal(lengths_to_mask(torch.from_numpy(g["mask_lengths"]), 6).numpy(), g["mask"])
    for row, (k, s, d, p) in zip(g["seq_len_out"], g["same_padding"]):
        assert get_same_padding(int(k), int(s), int(d)) == int(p)
        got = conv_out_length(torch.from_numpy(g["seq_len_in"]), int(k), int(s), int(p), int(d))
        assert np.array_equal(got.numpy(), row)
        assert [conv_out_length(int(v), int(k), int(s), int(p), int(d)) for v in g["seq_len_in"]] == list(row)
    with pytest.raises(ValueError):
        get_same_padding(3, 2, 2)


def test_encoders_load_reference_state_dicts_strictly():
    enc = QuartznetEncoder(repeat_blocks=1)
    st = synth.encoder_state(synth.quartznet_block_list(repeat_blocks=1), seed=5)
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    assert sum(p.numel() for p in enc.parameters()) == 6683456                 # SURVEY.md 8 a10 [probe]
    assert sum(p.numel() for p in QuartznetEncoder(repeat_blocks=3).parameters()) == 18894656
    keys = list(enc.state_dict().keys())
    assert "0.mconv.0.conv.weight" in keys and "0.mconv.1.conv.weight" in keys
    assert "0.mconv.2.layer.0.running_var" in keys and "1.res.0.conv.weight" in keys and "1.res.1.layer.0.bias" in keys
    c = synth.CITRINET_1024
    cn = CitrinetEncoder(c["filters"], c["kernel_sizes"], c["strides"])
    assert sum(p.numel() for p in cn.parameters()) == 139624336               # 139.6 M, SURVEY.md 8 a12
    assert "1.mconv.23.layer.0.fc.0.weight" in cn.state_dict() and "1.mconv.23.layer.0.fc.2.weight" in cn.state_dict()
    d = conv1d_decoder(1024, 29)
    assert list(d.state_dict().keys()) == ["weight", "bias"] and d.weight.shape == (29, 1024, 1)


def test_block_plan_folds_batchnorm_exactly():
    rng = np.random.Generator(np.random.PCG64(42))
    st = synth.block_state(rng, "", 16, 24, 3, 5, True, True, se=True)
    blk = CitrinetBlock(16, 24, repeat=3, kernel_size=(5,), stride=(2,), residual=True, separable=True).eval()
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    plan = build_block_plan(blk)
    assert [(s.K, s.S, s.D, s.P) for s in plan.subs] == [(5, 1, 1, 2), (5, 1, 1, 2), (5, 2, 1, 2)]   # stride on the last only
    assert plan.res_stride == 2 and plan.se_w1.shape == (3, 24) and plan.se_w2.shape == (24, 3)
    # folded pointwise == BN(conv1x1(x)) on a random vector (fp32 reference of the folding algebra)
    x = torch.randn(2, 24, 7)
    pw, bn = blk.mconv[6], blk.mconv[7].layer[0]
    ref = bn(pw.conv(x))
    sub = plan.subs[1]
    got = torch.einsum("oc,bct->bot", sub.pw_w.float(), x) + sub.shift[None, :, None]
    assert (got - ref).abs().max() < 2e-2 * ref.abs().max()                 # bf16 weights
    scale = bn.weight / torch.sqrt(bn.running_var + 1e-3)
    assert torch.allclose(sub.shift, bn.bias - bn.running_mean * scale, atol=1e-6)
    qb = QuartznetBlock(8, 8, repeat=2, kernel_size=(3,), stride=(2,), residual=True, separable=True).eval()
    assert build_block_plan(qb).res_stride == 4                               # stride ** repeat (quartznet/blocks.py:301)
    # non-separable kernel_size > 1 (the block's default): BN-folded [Cout, Cin, K] weight viewed as the [Cout, Cin * K]
    # operand of one GEMM over im2col rows (channel index c * K + k)
    fb = QuartznetBlock(8, 12, repeat=1, kernel_size=(11,), separable=False, residual=False).eval()
    fbn = fb.mconv[1].layer[0]
    fbn.running_var.uniform_(0.5, 1.5); fbn.weight.data.uniform_(0.5, 1.5); fbn.bias.data.normal_(0, 0.1)
    fplan = build_block_plan(fb)
    fs = fplan.subs[0]
    assert fs.full and fs.dw_w is None and tuple(fs.pw_w.shape) == (12, 8 * 11) and (fs.K, fs.S, fs.P) == (11, 1, 5)
    assert fplan.in_channels == 8 and fplan.out_channels == 12
    xin = torch.randn(2, 8, 30)
    cols = torch.nn.functional.unfold(xin[:, :, None, :], (1, 11), padding=(0, 5))          # [B, Cin * K, T], index c * K + k
    got = torch.einsum("ok,bkt->bot", fs.pw_w.float(), cols) + fs.shift[None, :, None]
    ref = fbn(fb.mconv[0].conv(xin))
    assert (got - ref).abs().max() < 2e-2 * ref.abs().max()
    with pytest.raises(NotImplementedError):
        bad = QuartznetBlock(8, 8, repeat=1, kernel_size=(3,), separable=False).eval()
        bad.mconv[0].conv.groups = 2          # grouped but not depthwise: not part of the reference's blocks
        build_block_plan(bad)


def test_greedy_decode_string_rules(golden_decode):
    tt = BatchTextTransformer(synth.quartznet_vocab())
    assert tt.vocab.blank_idx == 28 and tt.num_tokens == 29
    assert tt.decode_prediction(torch.from_numpy(golden_decode["ids"])) == list(golden_decode["text"])
    tb = BatchTextTransformer(synth.citrinet_vocab(64))
    assert tb.decode_prediction(torch.from_numpy(golden_decode["ids_bpe"])) == list(golden_decode["text_bpe"])
    # reference known answers (tests/text/test_transforms.py:59-91)
    v = BatchTextTransformer([" "] + [chr(c) for c in range(97, 123)], "<blank>", "<blank>", "<unk>", "<bos>", "<eos>")
    blank = torch.full((1, 100), v.vocab.blank_idx)
    assert v.decode_prediction(blank) == [""]
    x = blank.clone(); x[:, :10] = v.vocab.stoi["a"]; x[:, 15:20] = v.vocab.stoi["b"]
    assert v.decode_prediction(x) == ["ab"]
    x = blank.clone(); x[:, :10] = v.vocab.stoi["a"]; x[:, 15:20] = v.vocab.stoi["a"]
    assert v.decode_prediction(x) == ["aa"]
    assert Vocabulary(["a", "b", "c"]).blank_idx == 3                         # tests/text/test_vocab.py:108-112
    # collapsed-ids entry point used by predict(): same strings
    ids = golden_decode["ids"]
    rows = R.ctc_collapse(ids)
    col = np.full(ids.shape, -1, np.int64)
    for b, r in enumerate(rows):
        col[b, :len(r)] = r
    cnt = np.array([len(r) for r in rows], np.int32)
    assert tt.decode_collapsed(torch.from_numpy(col), torch.from_numpy(cnt)) == list(golden_decode["text"])


def test_torch_port_matches_reference_golden(golden_features, golden_e2e):
    """oracle/ref_torch.py (the cpu_baseline / --impl reference port) is pinned to the same golden vectors."""
    from oracle import ref_torch as RT

    x = synth.audio(3, 4800, 11, "noise")
    lens = synth.ragged_lengths(3, 4800, 111)
    f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens))
    assert np.abs(f.numpy() - golden_features["qn_noise.features"]).max() < 1e-4
    assert np.array_equal(fl.numpy(), golden_features["qn_noise.lengths"])
    cfgs = R.quartznet_cfgs(repeat_blocks=1)
    st = RT.to_torch(synth.encoder_state(synth.quartznet_block_list(repeat_blocks=1), seed=5))
    dec = RT.to_torch(synth.decoder_state(1024, 29, seed=6))
    ids, logits, el = RT.predict_ids(torch.from_numpy(synth.audio(2, 12000, 21, "tones")), cfgs, st, dec["weight"],
                                     dec["bias"], 64)
    ref = golden_e2e["qn5x5.full.logits"]
    assert np.abs(logits.numpy() - ref).max() < 1e-4 * np.abs(ref).max()
    assert np.array_equal(ids.numpy(), golden_e2e["qn5x5.full.ids"])


def test_shard_bounds_cover_batch_exactly():
    for n in (0, 1, 7, 128, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from thunder_speech_b200.parallel import sharded_predict, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
audio = torch.arange(7 * 4, dtype=torch.float32).reshape(7, 4)
calls = []
def fake_predict(x):                       # stands in for CTCModule.predict on this rank's shard
    calls.append(x.shape[0])
    return ["utt%d" % int(r[0].item() // 4) for r in x]
# training path: one flat gradient buffer, averaged in place over the ranks
from thunder_speech_b200.parallel import allreduce_mean_, flat_grad_views
lin = torch.nn.Linear(3, 2)
flat = flat_grad_views(lin.parameters())
assert flat.numel() == 8 and lin.weight.grad.data_ptr() == flat.data_ptr()
lin.weight.grad.fill_(1.0 + dist.get_rank()); lin.bias.grad.fill_(10.0 * dist.get_rank())
allreduce_mean_(flat)
assert torch.allclose(lin.weight.grad, torch.full((2, 3), 1.5)) and torch.allclose(lin.bias.grad, torch.full((2,), 5.0))
from thunder_speech_b200.parallel import ensure_grad_views
assert not ensure_grad_views(list(lin.parameters()), flat)
lin.bias.grad = None                        # what optimizer.zero_grad(set_to_none=True) does
assert ensure_grad_views(list(lin.parameters()), flat) and lin.bias.grad.data_ptr() == flat.data_ptr() + 6 * 4
out = sharded_predict(fake_predict, audio)
assert out == ["utt%d" % i for i in range(7)], out
lo, hi = shard_bounds(7, 2, dist.get_rank())
assert calls == [hi - lo]
# strong-scaling serving loop: every rank streams ITS slice of each batch, transcripts gathered once at the end
from thunder_speech_b200.parallel import sharded_predict_stream
class FakeModule:
    def __init__(self): self.seen = []
    def predict_stream(self, batches, depth=3):
        for xb in batches:
            self.seen.append(xb.shape[0])
            yield ["utt%d" % int(r[0].item() // 4) for r in xb]
fm = FakeModule()
batches = [audio, audio.flip(0), audio[:5]]
got = sharded_predict_stream(fm, iter(batches))
assert got == [["utt%d" % i for i in range(7)], ["utt%d" % i for i in reversed(range(7))], ["utt%d" % i for i in range(5)]], got
assert fm.seen == [hi - lo, hi - lo, shard_bounds(5, 2, dist.get_rank())[1] - shard_bounds(5, 2, dist.get_rank())[0]]
fm2 = FakeModule()
mine = [b[slice(*shard_bounds(b.shape[0], 2, dist.get_rank()))] for b in batches]
assert sharded_predict_stream(fm2, iter(mine), presharded=True) == got
# the overlapped gather falls back to ONE all_gather_object when a later batch outgrows the slot agreed on the first batch,
# or a transcript contains the separator: same result, nothing truncated; empty transcripts and empty shards survive
class TextModule:
    def __init__(self, rows): self.rows = rows
    def predict_stream(self, batches, depth=3):
        for xb, r in zip(batches, self.rows):
            yield list(r)
r = dist.get_rank()
rows = [["a%d" % r, ""], ["b" * (5000 + r), "c"], ["d", "e%d" % r]]
exp = [["a0", "", "a1", ""], ["b" * 5000, "c", "b" * 5001, "c"], ["d", "e0", "d", "e1"]]
assert sharded_predict_stream(TextModule(rows), iter([audio] * 3), presharded=True) == exp
rows = [["x", "y\x00z"] if r == 1 else ["x", "y"], [] if r == 0 else ["only%d" % r]]
assert sharded_predict_stream(TextModule(rows), iter([audio] * 2), presharded=True) == [["x", "y", "x", "y\x00z"], ["only1"]]
rows = [["é字 %d" % r, ""], [], ["", ""]]
assert sharded_predict_stream(TextModule(rows), iter([audio] * 3), presharded=True) == [["é字 0", "", "é字 1", ""], [], ["", "", "", ""]]
dist.barrier(); dist.destroy_process_group()
print("rank", sys.argv[3], "ok")
"""


def test_sharded_predict_world_size_2_gloo(tmp_path):
    """N > 1 paths on CPU: two gloo ranks shard a batch with no data-path collective and gather the transcripts; the
    training step's flat gradient buffer is averaged in place."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = str(s.getsockname()[1]); s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_vectorised_detokenisation_equals_reference_rules():
    """`_finish` (code-point table / object-array join) == the reference's per-token join + replacements + special-token
    removal (text_processing/transform.py:93-122, vocab.py:85-130) on random id rows, for a character vocabulary with
    all special tokens, a word-piece vocabulary, and a vocabulary whose ordinary tokens spell a special token."""
    rng = np.random.default_rng(5)

    def slow(tt, row):
        out = "".join(tt.vocab.itos[int(i)] for i in row)
        out = out.replace("▁", " ").replace("|", " ")
        return tt.vocab.remove_special_tokens(out)

    cases = [
        BatchTextTransformer(synth.quartznet_vocab(), pad_token="<pad>", unknown_token="<unk>", start_token="<bos>",
                             end_token="<eos>"),
        BatchTextTransformer(synth.citrinet_vocab(256)),
        BatchTextTransformer(["<", "b", "l", "a", "n", "k", ">", "|", "▁", "é", "字"]),   # "<blank>" can be spelled
    ]
    for tt in cases:
        V = len(tt.vocab.itos)
        for _ in range(50):
            row = rng.integers(0, V, rng.integers(0, 60))
            assert tt._finish(row) == slow(tt, row)
        spelled = [tt.vocab.stoi[c] for c in "<blank>" if c in tt.vocab.stoi]
        if len(spelled) == 7:
            assert tt._finish(np.array(spelled + [tt.vocab.stoi["a"]])) == "a"      # removed as a substring, like the reference
    col = torch.from_numpy(rng.integers(0, 29, (4, 12)).astype(np.int64))
    cnt = torch.tensor([12, 0, 5, 1], dtype=torch.int32)
    tt = BatchTextTransformer(synth.quartznet_vocab())
    assert tt.decode_collapsed(col, cnt) == [slow(tt, col[b, :cnt[b]].numpy()) for b in range(4)]


def test_text_encode_known_answers_and_tokenizers(tmp_path):
    """BatchTextTransformer.encode with the reference's own known answer (tests/text/test_transforms.py:41-57), the
    no-unknown-token filtering rule (vocab.py:77-80), a custom tokenizer and -- when the reference's sample model is around
    -- sentencepiece pieces."""
    from string import ascii_lowercase

    from thunder_speech_b200.text_processing import word_tokenizer

    tfm = BatchTextTransformer(tokens=[" "] + list(ascii_lowercase), blank_token="<blank>", pad_token="<blank>",
                               unknown_token="<unk>", start_token="<bos>", end_token="<eos>")
    encoded, lens = tfm.encode(["hello world", "oi"], return_length=True)
    assert encoded.shape == (2, 13) and encoded.dtype == torch.long
    assert encoded[0].tolist() == [29, 8, 5, 12, 12, 15, 0, 23, 15, 18, 12, 4, 30]
    assert lens.tolist() == [13, 4]
    assert encoded[1].tolist() == [29, 15, 9, 30] + [tfm.vocab.pad_idx] * 9
    assert torch.equal(tfm.encode(["hello world", "oi"], return_length=False), encoded)
    assert tfm.encode(["a!b"])[0].tolist() == [[29, 1, 28, 2, 30]]                     # unknown -> <unk>
    plain = BatchTextTransformer(synth.quartznet_vocab())
    y, yl = plain.encode(["ab!c", ""])
    assert y.tolist() == [[1, 2, 3], [28, 28, 28]] and yl.tolist() == [3, 0]            # "!" dropped, pad = blank
    words = BatchTextTransformer(["hello", "world"], custom_tokenizer_function=word_tokenizer)
    assert words.encode(["hello there world"])[0].tolist() == [[0, 1]]
    sp = "/root/reference/tests/nemo_config_samples/example_tokenizer.model"
    if os.path.exists(sp):
        bpe = BatchTextTransformer(["▁the", "s", "t", "▁ca"], sentencepiece_model=sp)
        assert bpe.tokenizer("the cats") == ["▁the", "▁ca", "t", "s"]
        assert bpe.encode(["the cats"])[0].tolist() == [[0, 3, 2, 1]]
        # from_sentencepiece (text_processing/transform.py:124-150): tokenizer.vocab lines "piece<TAB>score", specials skipped
        import shutil

        shutil.copy(sp, tmp_path / "tokenizer.model")
        (tmp_path / "tokenizer.vocab").write_text("<unk>\t0\n<s>\t0\n</s>\t0\n▁the\t-1.5\ns\t-2\nt\t-2.5\n▁ca\t-3\n",
                                                  encoding="utf-8")
        fsp = BatchTextTransformer.from_sentencepiece(str(tmp_path))
        assert list(fsp.vocab.itos[:4]) == ["▁the", "s", "t", "▁ca"] and fsp.encode(["the cats"])[0].tolist() == [[0, 3, 2, 1]]


def test_product_code_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under the package imports it, and ``bench.py`` only does so inside
    its CPU-baseline function.  Nothing that runs on the GPU box reads /root/reference."""
    pkg = os.path.join(ROOT, "thunder_speech_b200")
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                assert not pat.search(src), f"{fn} imports the oracle"
                assert "/root/reference" not in src, f"{fn} reads the reference tree"
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    hits = [m.start() for m in pat.finditer(bench)]
    assert hits, "bench.py must time the oracle port as its cpu_baseline"
    start = bench.index("def cpu_reference_rate")
    end = bench.index("\ndef ", start + 1)
    assert all(start < h < end for h in hits), "oracle imported outside bench.py cpu_reference_rate()"
    assert "/root/reference" not in bench
    # the unmodified reference (baseline/_ref, git-ignored) is only ever reached through baseline/ref_harness.py, and
    # only from bench.py's reference arms
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                assert "ref_harness" not in src and "baseline" not in src.replace("cpu_baseline", ""), fn


def test_error_rate_metrics_match_levenshtein_definition():
    """thunder_speech_b200.metrics (stand-in for torchmetrics' CharErrorRate / WordErrorRate used by validation_step,
    src/thunder/module.py:17-18,157-162): vectorised edit distance == the textbook DP; rates accumulate as
    total errors / total target length."""
    import random

    from thunder_speech_b200.metrics import CharErrorRate, WordErrorRate, edit_distance

    def dp(a, b):
        d = list(range(len(b) + 1))
        for i, x in enumerate(a, 1):
            nd = [i]
            for j, yv in enumerate(b, 1):
                nd.append(min(d[j] + 1, nd[j - 1] + 1, d[j - 1] + (x != yv)))
            d = nd
        return d[-1]

    rnd = random.Random(0)
    for _ in range(500):
        a = "".join(rnd.choice("abc ") for _ in range(rnd.randint(0, 14)))
        b = "".join(rnd.choice("abc ") for _ in range(rnd.randint(0, 14)))
        assert edit_distance(a, b) == dp(a, b)
        assert edit_distance(a.split(), b.split()) == dp(a.split(), b.split())
    cer, wer = CharErrorRate(), WordErrorRate()
    assert cer(["abc", "hello"], ["abd", "hallo"]) == pytest.approx(2 / 8)
    assert cer("", "ab") == pytest.approx(1.0)
    assert cer.compute() == pytest.approx(4 / 10)
    assert wer(["the cat sat"], ["the cat sat down"]) == pytest.approx(0.25)
    assert wer(["a b c d"], ["a x c"]) == pytest.approx(2 / 3)
    assert wer.compute() == pytest.approx(3 / 7)
    wer.reset()
    assert np.isnan(wer.compute())
    with pytest.raises(ValueError):
        cer(["a"], ["a", "b"])


def test_ctc_module_constructor_and_configure_optimizers_like_the_reference():
    """BaseCTCModule's constructor arguments and configure_optimizers (src/thunder/module.py:26-64,165-192)."""
    import types

    from thunder_speech_b200.module import BaseCTCModule

    def make(**kw):
        enc = QuartznetEncoder(filters=[16, 16, 16, 16, 16], kernel_sizes=[3, 3, 3, 3, 3], repeat_blocks=1)
        return BaseCTCModule(enc, conv1d_decoder(1024, 29), FilterbankFeatures(), BatchTextTransformer(synth.quartznet_vocab()), **kw)

    m = make()
    opt = m.configure_optimizers()
    assert isinstance(opt, torch.optim.AdamW)
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == sum(p.numel() for p in m.parameters())
    m = make(optimizer_class=torch.optim.SGD, optimizer_kwargs={"lr": 0.1, "momentum": 0.9},
             lr_scheduler_class=torch.optim.lr_scheduler.OneCycleLR,
             lr_scheduler_kwargs={"max_lr": 0.5, "total_steps_arg": "total_steps", "interval": "epoch"},
             encoder_final_dimension=1024)
    with pytest.raises(RuntimeError):
        m.configure_optimizers()                      # total_steps_arg needs the trainer
    m.trainer = types.SimpleNamespace(estimated_stepping_batches=123)
    cfg = m.configure_optimizers()
    assert isinstance(cfg["optimizer"], torch.optim.SGD) and cfg["optimizer"].defaults["momentum"] == 0.9
    assert cfg["lr_scheduler"]["interval"] == "epoch" and cfg["lr_scheduler"]["scheduler"].total_steps == 123
    assert m.encoder_final_dimension == 1024 and "total_steps_arg" in m.lr_scheduler_kwargs   # kwargs left intact


def test_every_entry_point_is_documented_against_the_reference():
    """INTEGRATION.md maps each C-ABI entry point to the reference interface it replaces; the header cites reference lines."""
    hdr = open(os.path.join(ROOT, "include", "thunder_b200.h"), encoding="utf-8").read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()
    syms = set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", hdr))
    assert len(syms) >= 40
    missing = sorted(s for s in syms if s not in doc)
    assert not missing, missing
    assert len(re.findall(r"[a-z_/]+\.py:\d+", hdr)) >= 30      # reference file:line citations in the header
