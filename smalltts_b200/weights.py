"""Weight readers for the engine: every file format the reference's assets come in, without onnxruntime / onnx.

The reference ships its models as four ONNX graphs fetched from the Hub (assets/ensure.py:33-40:
``dmd/condition_encoder.onnx``, ``dmd/denoiser.onnx``, ``codec/decoder.onnx``, ``codec/encoder.onnx``) and trains /
distils them as ``.pt`` state dicts (scripts/train/dmd2/distill.py:468-479).  The engine wants fp32 tensors under the
PyTorch state-dict names (include/smalltts_b200.h, ``stts_load_weight``), so this module turns any of

* ``.pt`` / ``.pth`` / ``.ckpt``  (torch.save of a state dict, possibly wrapped: ``{"model": ...}``, EMA / DDP /
  ``torch.compile`` prefixes -- distill.py:39-57),
* ``.safetensors``,
* ``.onnx``  (the initialisers and folded constants of an exported graph; parsed here with a ~100-line protobuf
  wire-format reader because neither ``onnx`` nor ``onnxruntime`` is a dependency of this package),
* ``.sttsw``  (this package's own flat container, written by :func:`save_packed` -- one mmap-able file per model)

into ``{state_dict_name: np.ndarray}`` and checks the result against the architecture's tensor list
(:func:`smalltts_b200.synthetic.dit_specs` etc.) so that a missing or mis-shaped tensor is reported by name
before anything is uploaded.

ONNX naming.  Exporters keep ``nn.Parameter`` names for most initialisers (``dit.blocks.0.attn.to_q.bias``), but
the TorchScript exporter folds ``Linear`` weights into anonymous, already transposed ``onnx::MatMul_1234`` tensors
and some Conv weights into ``onnx::Conv_987``.  Those are resolved through the graph: the consuming node's name
(``/dit/blocks.0/attn/to_q/MatMul``) spells the module path, which gives ``dit.blocks.0.attn.to_q.weight``.  Names
are finally matched to the expected list by longest unique suffix, so wrapper prefixes (``model.``, ``m.``) do not
matter.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------------------------
# protobuf wire format (just enough for onnx.ModelProto; field numbers from onnx/onnx.proto3)
# ----------------------------------------------------------------------------------------------------------------


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) of one message; length-delimited values are memoryview slices."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos : pos + 8]), pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > end:
                raise ValueError("truncated protobuf message")
            v, pos = buf[pos : pos + n], pos + n
        elif wt == 5:
            v, pos = bytes(buf[pos : pos + 4]), pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed64(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


# onnx.TensorProto.DataType -> numpy
_ONNX_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 4: np.uint16, 5: np.int16, 6: np.int32, 7: np.int64,
                9: np.bool_, 10: np.float16, 11: np.float64, 12: np.uint32, 13: np.uint64}
_BFLOAT16 = 16


def _bf16_to_f32(raw: np.ndarray) -> np.ndarray:
    return (raw.astype(np.uint32) << 16).view(np.float32)


def _parse_tensor(buf: memoryview, base_dir: str) -> Tuple[str, Optional[np.ndarray]]:
    """onnx.TensorProto -> (name, array).  Handles raw_data, the typed repeated fields and external data."""
    dims: List[int] = []
    dtype = 0
    name = ""
    raw = None
    floats: List[np.ndarray] = []
    ints: List[int] = []
    doubles: List[np.ndarray] = []
    external: Dict[str, str] = {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += [_signed64(x) for x in _packed_varints(v, wt)]
        elif fno == 2:
            dtype = v
        elif fno == 4:  # float_data
            floats.append(np.frombuffer(v, dtype="<f4"))
        elif fno in (5, 7, 11):  # int32_data / int64_data / uint64_data
            ints += _packed_varints(v, wt)
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = v
        elif fno == 10:
            doubles.append(np.frombuffer(v, dtype="<f8"))
        elif fno == 13:  # external_data: StringStringEntryProto
            kv = {f: bytes(x).decode() for f, _, x in _fields(v)}
            external[kv.get(1, "")] = kv.get(2, "")
    if external:
        loc = os.path.join(base_dir, external["location"])
        off = int(external.get("offset", 0))
        n = external.get("length")
        with open(loc, "rb") as fh:
            fh.seek(off)
            raw = memoryview(fh.read(int(n)) if n is not None else fh.read())
    count = int(np.prod(dims)) if dims else 1
    if dtype == _BFLOAT16:
        if raw is not None:
            arr = _bf16_to_f32(np.frombuffer(raw, dtype="<u2"))
        else:
            arr = _bf16_to_f32(np.asarray(ints, dtype=np.uint16))
    elif dtype in _ONNX_DTYPES:
        np_dt = np.dtype(_ONNX_DTYPES[dtype])
        if raw is not None:
            arr = np.frombuffer(raw, dtype=np_dt.newbyteorder("<"))
        elif floats:
            arr = np.concatenate(floats)
        elif doubles:
            arr = np.concatenate(doubles)
        elif dtype == 10:  # fp16 bit patterns travel in int32_data
            arr = np.asarray(ints, dtype=np.uint16).view(np.float16)
        else:
            arr = np.asarray([_signed64(x) for x in ints], dtype=np.int64).astype(np_dt)
    else:
        return name, None  # strings, complex, fp8: never a weight of this model
    if arr.size != count:
        raise ValueError(f"ONNX tensor {name!r}: {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


class OnnxGraph:
    """The parts of an ONNX file the loader needs: initialisers / Constant tensors and node wiring."""

    def __init__(self) -> None:
        self.tensors: Dict[str, np.ndarray] = {}
        self.nodes: List[Tuple[str, str, List[str], List[str]]] = []  # (name, op_type, inputs, outputs)
        self.inputs: List[str] = []
        self.outputs: List[str] = []


def read_onnx(path: str) -> OnnxGraph:
    """Parse ``path`` (ModelProto.graph = 7; GraphProto.node = 1, initializer = 5, input = 11, output = 12;
    NodeProto.input = 1, output = 2, name = 3, op_type = 4, attribute = 5; AttributeProto.name = 1, t = 5)."""
    base = os.path.dirname(os.path.abspath(path))
    data = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
    g = OnnxGraph()
    graph_buf = None
    for fno, wt, v in _fields(memoryview(data)):
        if fno == 7 and wt == 2:
            graph_buf = v
    if graph_buf is None:
        raise ValueError(f"{path}: no graph in this file (not an ONNX ModelProto?)")
    for fno, wt, v in _fields(graph_buf):
        if fno == 5:
            name, arr = _parse_tensor(v, base)
            if arr is not None:
                g.tensors[name] = arr
        elif fno == 1:
            name = op = ""
            ins: List[str] = []
            outs: List[str] = []
            const = None
            for f2, _, x in _fields(v):
                if f2 == 1:
                    ins.append(bytes(x).decode())
                elif f2 == 2:
                    outs.append(bytes(x).decode())
                elif f2 == 3:
                    name = bytes(x).decode()
                elif f2 == 4:
                    op = bytes(x).decode()
                elif f2 == 5:
                    aname, at = "", None
                    for f3, _, y in _fields(x):
                        if f3 == 1:
                            aname = bytes(y).decode()
                        elif f3 == 5:
                            at = y
                    if aname == "value" and at is not None:
                        const = at
            g.nodes.append((name, op, ins, outs))
            if op == "Constant" and const is not None and outs:
                _, arr = _parse_tensor(const, base)
                if arr is not None and arr.size > 1:
                    g.tensors[outs[0]] = arr
        elif fno in (11, 12):
            for f2, _, x in _fields(v):
                if f2 == 1:
                    (g.inputs if fno == 11 else g.outputs).append(bytes(x).decode())
    return g


def _module_path(node_name: str) -> str:
    """'/dit/blocks.0/attn/to_q/MatMul' -> 'dit.blocks.0.attn.to_q' (TorchScript exporter scope names)."""
    parts = [p for p in node_name.split("/") if p]
    return ".".join(parts[:-1]) if len(parts) > 1 else ""


def onnx_state_dict(path: str, specs: Optional[Iterable[Tuple]] = None, exec_rank=None) -> Dict[str, np.ndarray]:
    """Initialisers of an exported graph under (approximately) their PyTorch names.

    Named initialisers are kept as they are.  Anonymous ``onnx::MatMul_*`` tensors become
    ``<module path of the MatMul node>.weight`` transposed back to nn.Linear's ``[out, in]``; anonymous
    ``onnx::Conv_*`` / ``onnx::ConvTranspose_*`` tensors become ``<module path>.weight`` (input 1) or ``.bias``
    (input 2); a Gather's anonymous table becomes ``<module path>.weight`` (nn.Embedding).

    Bare ``nn.Parameter`` members used in arithmetic (the ConvNeXt layer scales ``gamma.unsqueeze(-1)``,
    modeling_vibevoice_acoustic_tokenizer.py:289,296) are constant-folded into anonymous ``onnx::Mul_*`` tensors
    whose consumer is named after the *owning* module only (``/decoder/stem/stage.0/Mul_1``).  With ``specs`` (the
    architecture's (name, shape, ...) list) they are resolved too: the anonymous operands of elementwise nodes in
    module scope M are assigned, in execution order, to the still-unmatched direct parameters ``M.<leaf>`` of the
    same element count, in declaration order.  One parameter is not used as it is: ``style_encoder.log_scale`` enters
    as ``x * exp(log_scale)`` (style.py:167), so a folded graph holds ``exp(log_scale)`` as the Mul operand; the log is
    taken back.

    Graphs traced through direct method calls (``DiTModel.denoise_step`` calls ``self.dit.forward_cached(...)``, not
    ``self.dit(...)``, models/backbone/model.py:97-100) lose the outer scopes: their nodes are called
    ``/to_q/MatMul``, ``/to_q_1/MatMul`` ... and the node name no longer identifies the layer.  Two more rules cover
    that (checked on a TorchScript export of the reference's real denoiser, oracle/make_golden_onnx.py --full):
      * bias pairing -- ``MatMul(x, W)`` whose output feeds ``Add(<named> X.bias, .)`` is ``X.weight``;
      * execution order -- what is still anonymous (bias-free nn.Linear) is assigned, in node order, to the first
        unmatched 2-D tensor of ``specs`` with that shape, ``specs`` being sorted by ``exec_rank(name)`` (the order in
        which the reference's forward runs those layers; declaration order if no ranking is given)."""
    g = read_onnx(path)
    out: Dict[str, np.ndarray] = {}
    anonymous = {n for n in g.tensors if n.startswith("onnx::") or n.startswith("/") or n.isdigit()}
    for n, a in g.tensors.items():
        if n not in anonymous:
            out[n] = a
    used = set()
    for name, op, ins, _ in g.nodes:
        mod = _module_path(name)
        if not mod:
            continue
        for idx, src in enumerate(ins):
            if src not in anonymous:
                continue
            a = g.tensors[src]
            key = None
            if op == "MatMul" and idx == 1 and a.ndim == 2:
                key, a = mod + ".weight", np.ascontiguousarray(a.T)
            elif op == "Gemm" and idx in (1, 2):
                key = mod + (".weight" if idx == 1 else ".bias")
            elif op in ("Conv", "ConvTranspose") and idx in (1, 2):
                key = mod + (".weight" if idx == 1 else ".bias")
            elif op == "Gather" and idx == 0 and a.ndim == 2:
                key = mod + ".weight"
            if key is not None:
                out.setdefault(key, a)
                used.add(src)
    # ---- bias pairing (and detection of node names that do not spell a full module path)
    consumers: Dict[str, List[int]] = {}
    for i, (_, _, ins, _) in enumerate(g.nodes):
        for src in ins:
            consumers.setdefault(src, []).append(i)
    unresolved: List[Tuple[str, np.ndarray]] = []  # anonymous MatMul weights in execution order
    for name, op, ins, outs in g.nodes:
        if op != "MatMul" or len(ins) < 2 or ins[1] not in anonymous or g.tensors[ins[1]].ndim != 2:
            continue
        w = g.tensors[ins[1]]
        paired = None
        for ci in consumers.get(outs[0], []) if outs else []:
            _, cop, cins, _ = g.nodes[ci]
            if cop == "Add":
                for other in cins:
                    if other in g.tensors and other not in anonymous and other.endswith(".bias") \
                            and g.tensors[other].shape == (w.shape[1],):
                        paired = other[: -len("bias")] + "weight"
        if paired is not None:
            out[paired] = np.ascontiguousarray(w.T)
            stale = _module_path(name) + ".weight"
            if stale != paired and out.get(stale) is not None and out[stale].shape == w.T.shape \
                    and np.array_equal(out[stale], w.T):
                del out[stale]  # the scope-derived guess was only a fragment of the real name
        else:
            unresolved.append((_module_path(name) + ".weight", w))
    if specs is not None:
        specs = list(specs)
        have0 = lambda name: any(k == name or k.endswith("." + name) for k in out)  # noqa: E731
        names = {sp[0] for sp in specs}
        todo = [(sp[0], tuple(sp[1])) for sp in specs if len(sp[1]) == 2 and not have0(sp[0])
                and sp[0].endswith(".weight") and sp[0][: -len("weight")] + "bias" not in names]
        if exec_rank is not None:
            todo.sort(key=lambda it: exec_rank(it[0]))
        for guess, w in unresolved:
            hits = [k for k, _ in todo if k == guess or guess.endswith("." + k) or k.endswith("." + guess)]
            if len(hits) == 1 and sum(1 for n2, _, _, _ in g.nodes if _module_path(n2) + ".weight" == guess) == 1:
                todo = [it for it in todo if it[0] != hits[0]]
                continue  # the node name is a unique module path naming exactly one expected tensor (rule 1)
            for i, (pname, shape) in enumerate(todo):
                if shape == w.shape[::-1]:
                    out[pname] = np.ascontiguousarray(w.T)
                    if guess in out and guess != pname:
                        del out[guess]
                    del todo[i]
                    break
    if specs is not None:
        have = lambda name: any(k == name or k.endswith("." + name) for k in out)  # noqa: E731
        pending: Dict[str, List[Tuple[str, Tuple[int, ...]]]] = {}
        for spec in specs:
            name, shape = spec[0], tuple(spec[1])
            if "." in name and not have(name):
                pending.setdefault(name.rsplit(".", 1)[0], []).append((name, shape))
        for name, op, ins, _ in g.nodes:
            if op not in ("Mul", "Add", "Sub", "Div"):
                continue
            parts = [p for p in name.split("/") if p]
            mod = ".".join(parts[:-1])
            owner = next((m for m in pending if mod == m or mod.endswith("." + m)), None)
            if owner is None:
                continue
            for src in ins:
                if src in anonymous and src not in used:
                    a = g.tensors[src]
                    for i, (pname, shape) in enumerate(pending[owner]):
                        if int(np.prod(shape, dtype=np.int64)) == a.size:
                            if pname.endswith("log_scale") and op == "Mul":
                                # style.py:167 multiplies by exp(log_scale); a folding exporter leaves exp(.) behind
                                a = np.log(np.asarray(a, dtype=np.float64)).astype(np.float32)
                            out[pname] = a.reshape(shape)
                            used.add(src)
                            del pending[owner][i]
                            break
    return out


# ----------------------------------------------------------------------------------------------------------------
# state-dict files
# ----------------------------------------------------------------------------------------------------------------

_PREFIXES = ("ema_model.", "module.", "_orig_mod.", "online_model.")  # scripts/train/dmd2/distill.py:39-57
_EXTRAS = ("initted", "step")  # ema_pytorch bookkeeping saved next to the weights (distill.py:468-470)
_CONTAINERS = ("student_model", "model", "state_dict")


def strip_prefixes(sd: Dict) -> Dict:
    out = {}
    for k, v in sd.items():
        changed = True
        while changed:
            changed = False
            for p in _PREFIXES:
                if k.startswith(p):
                    k, changed = k[len(p):], True
        if k in _EXTRAS:
            continue
        out[k] = v
    return out


def load_state_dict_file(path: str, container_keys: Sequence[str] = _CONTAINERS,
                         specs: Optional[Iterable[Tuple]] = None, exec_rank=None) -> Dict:
    """Read a checkpoint into {name: tensor-or-array}; strips the wrapper prefixes the reference's trainers leave
    behind and drops their ``initted`` / ``step`` extras (distill.py:39-57,468-470).  Format by extension."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".safetensors":
        from safetensors.numpy import load_file

        sd = load_file(path)
    elif ext == ".onnx":
        sd = onnx_state_dict(path, specs, exec_rank)
    elif ext == ".sttsw":
        sd = load_packed(path)
    else:
        import torch

        sd = torch.load(path, map_location="cpu", weights_only=True)
        for k in container_keys:
            if isinstance(sd, dict) and k in sd and isinstance(sd[k], dict):
                sd = sd[k]
                break
    return strip_prefixes(sd)


def vibevoice_native_to_hf(sd: Dict) -> Dict:
    """Rename microsoft/VibeVoice's own acoustic-tokenizer keys (``upsample_layers`` / ``downsample_layers`` /
    ``stages``, as in the ``model.acoustic_tokenizer.*`` tensors of microsoft/VibeVoice-1.5B and in graphs exported
    from that code) to the HF ``transformers`` names the engine uses (``stem`` / ``conv_layers.{i}.stage``).

    Best effort: written from the published module layout, not checked against the real files (none are available
    offline); :func:`match_to_specs` still validates every name and shape afterwards."""
    import re

    if not any("upsample_layers." in k or "downsample_layers." in k for k in sd):
        return sd
    out = {}
    for k, v in sd.items():
        n = k
        n = re.sub(r"(up|down)sample_layers\.0\.0\.", "stem.", n)
        n = re.sub(r"(up|down)sample_layers\.(\d+)\.0\.", lambda m: f"conv_layers.{int(m.group(2)) - 1}.", n)
        n = re.sub(r"stages\.0\.(\d+)\.", r"stem.stage.\1.", n)
        n = re.sub(r"stages\.(\d+)\.(\d+)\.", lambda m: f"conv_layers.{int(m.group(1)) - 1}.stage.{m.group(2)}.", n)
        n = n.replace("mixer.conv.conv.conv.", "mixer.conv.").replace("head.conv.conv.", "head.conv.")
        out[n] = v
    return out


def _shape(t) -> Tuple[int, ...]:
    return tuple(int(x) for x in t.shape)


def match_to_specs(sd: Dict, specs: Iterable[Tuple], what: str, strict: bool = True) -> Dict:
    """Pick the tensors of ``specs`` ((name, shape, ...) tuples) out of ``sd``.

    A key matches an expected name if it is equal to it or ends with ``"." + name`` (exporter / wrapper
    prefixes); the shape must agree (a 2-D tensor stored transposed, as the ONNX exporters do for MatMul, is
    turned back).  With ``strict`` every expected tensor must be found exactly once; the error lists what is
    missing, by name -- the engine's own ``stts_finalize_weights`` check would only say so after the upload."""
    by_suffix: Dict[str, List[str]] = {}
    for k in sd:
        parts = k.split(".")
        for i in range(len(parts)):
            by_suffix.setdefault(".".join(parts[i:]), []).append(k)
    out, missing, bad = {}, [], []
    for spec in specs:
        name, shape = spec[0], tuple(spec[1])
        cands = by_suffix.get(name, [])
        if name in sd:
            cands = [name]
        if len(cands) != 1:
            missing.append(name if not cands else f"{name} (ambiguous: {cands[:3]})")
            continue
        t = sd[cands[0]]
        if _shape(t) == shape:
            out[name] = t
        elif len(shape) == 2 and _shape(t) == shape[::-1]:
            out[name] = t.T
        elif int(np.prod(_shape(t), dtype=np.int64)) == int(np.prod(shape, dtype=np.int64)) and len(shape) <= 1:
            out[name] = t.reshape(shape)
        else:
            bad.append(f"{name}: expected {shape}, file has {_shape(t)}")
    if strict and (missing or bad):
        lines = [f"{what}: checkpoint does not match the architecture"]
        if missing:
            lines.append(f"  missing {len(missing)} tensors, e.g. {missing[:8]}")
        if bad:
            lines.append(f"  mis-shaped {len(bad)} tensors, e.g. {bad[:8]}")
        raise KeyError("\n".join(lines))
    return out


_ENC_LEAVES = ("wq", "wk", "wv", "gate", "wo", "w1", "w3", "w2")  # style.py:47-66,76-77 == phonemes.py blocks
_DIT_LEAVES = ("to_q", "to_k_self", "to_v_self", "to_k_ref", "to_v_ref", "to_k_text", "to_v_text", "gate", "to_out",
               "w1", "w3", "w2")  # dit.py:95-119,176-186


def dit_exec_rank(name: str) -> Tuple[int, int, int]:
    """Order in which DiTModel's forward runs its nn.Linear layers: style encoder, text encoder, DiT
    (models/backbone/model.py:88-100); inside a block the order of style.py:47-66 / dit.py:95-119."""
    import re

    group = 0 if name.startswith("style_encoder.") else 1 if name.startswith("phoneme_embedding.") else 2
    m = re.search(r"(?:blocks|transformer_blocks)\.(\d+)\.", name)
    block = int(m.group(1)) if m else (-1 if ".in_proj." in name or ".input_embed." in name else 10_000)
    parts = name.split(".")
    leaves = _DIT_LEAVES if group == 2 else _ENC_LEAVES
    leaf = next((leaves.index(p) for p in parts if p in leaves), len(leaves))
    return group, block, leaf


def load_model_weights(paths: Sequence[str], specs: Iterable[Tuple], what: str, exec_rank=None) -> Dict:
    """Merge one or more files (the reference splits the DiT over condition_encoder.onnx + denoiser.onnx,
    infer/onnx.py:60-62) and select the architecture's tensors."""
    specs = list(specs)
    merged: Dict = {}
    seen = set()
    for p in paths:
        ap = os.path.abspath(p)
        if ap in seen:
            continue
        seen.add(ap)
        merged.update(load_state_dict_file(p, specs=specs, exec_rank=exec_rank))
    merged = vibevoice_native_to_hf(merged)
    return match_to_specs(merged, specs, what)


# ----------------------------------------------------------------------------------------------------------------
# .sttsw: flat container (JSON index + 64-byte aligned little-endian payloads), readable with one mmap
# ----------------------------------------------------------------------------------------------------------------

_MAGIC = b"STTSW001"


def save_packed(path: str, sd: Dict, dtype: str = "float32") -> None:
    """Write {name: tensor} as one file: magic, u64 index length, JSON index, aligned tensor payloads.
    ``dtype``: "float32" (exact) or "bfloat16" (halves the file; the engine rounds GEMM weights to bf16 anyway,
    but norm scales / biases / adaLN tables stay fp32 on the device, so those are always written as fp32)."""
    if dtype not in ("float32", "bfloat16"):
        raise ValueError("dtype must be 'float32' or 'bfloat16'")
    index, blobs, off = {}, [], 0
    for name, t in sd.items():
        a = t.detach().cpu().numpy() if type(t).__module__.startswith("torch") else np.asarray(t)
        a = np.require(a, dtype=np.float32, requirements=["C"])  # keeps 0-d tensors 0-d
        use_bf16 = dtype == "bfloat16" and a.ndim >= 2
        if use_bf16:
            u = a.view(np.uint32)
            payload = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype("<u2").tobytes()  # round to nearest even
        else:
            payload = a.astype("<f4").tobytes()
        index[name] = {"shape": list(a.shape), "dtype": "bfloat16" if use_bf16 else "float32", "offset": off,
                       "nbytes": len(payload)}
        pad = (-len(payload)) % 64
        blobs.append(payload + b"\0" * pad)
        off += len(payload) + pad
    head = json.dumps(index).encode()
    head += b" " * ((-(len(_MAGIC) + 8 + len(head))) % 64)
    with open(path, "wb") as fh:
        fh.write(_MAGIC)
        fh.write(struct.pack("<Q", len(head)))
        fh.write(head)
        for b in blobs:
            fh.write(b)


def load_packed(path: str) -> Dict[str, np.ndarray]:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if bytes(mm[:8]) != _MAGIC:
        raise ValueError(f"{path}: not a .sttsw file")
    (n,) = struct.unpack("<Q", bytes(mm[8:16]))
    index = json.loads(bytes(mm[16 : 16 + n]).decode())
    base = 16 + n
    out = {}
    for name, m in index.items():
        raw = mm[base + m["offset"] : base + m["offset"] + m["nbytes"]]
        if m["dtype"] == "bfloat16":
            a = _bf16_to_f32(raw.view("<u2"))
        else:
            a = raw.view("<f4")
        out[name] = a.reshape(m["shape"])
    return out
