"""C++ host runtime (libodbhost.so): safetensors reader and the prefetching synthetic-token generator."""
import os

import pytest
import torch

from opendiloco_b200 import _lib
from opendiloco_b200.utils import safetensors_io as st
from opendiloco_b200.utils.data import NativeTokenLoader, SyntheticTokenLoader, collate_causal_lm, FakeTokenizedDataset

needs_host = pytest.mark.skipif(_lib.host_lib() is None, reason="libodbhost.so not built")


def test_safetensors_roundtrip(tmp_path):
    t = {"a.weight": torch.randn(7, 5), "b": torch.arange(12, dtype=torch.int64).reshape(3, 4), "c": torch.randn(9).to(torch.bfloat16)}
    p = str(tmp_path / "x.safetensors")
    st.save_safetensors(t, p, metadata={"format": "pt"})
    back = st.load_safetensors(p)
    assert set(back) == set(t) and all(torch.equal(back[k], t[k]) for k in t)
    hf = pytest.importorskip("safetensors.torch")
    ref = hf.load_file(p)                       # the file is a valid safetensors file for the reference's loader too
    assert all(torch.equal(ref[k], t[k]) for k in t)


@needs_host
def test_native_safetensors_matches_python_and_validates(tmp_path, ref_model_dir):
    if ref_model_dir is not None:
        path = os.path.join(ref_model_dir, "model.safetensors")
        a = st._load_native(path)
        header, base = st.read_header(path)
        assert len(a) == 21 and set(a) == {k for k in header if k != "__metadata__"}
    bad = tmp_path / "bad.safetensors"
    bad.write_bytes(b"\x10\x00\x00\x00\x00\x00\x00\x00" + b'{"x":{"dtype":"F32","shape":[4],"data_offsets":[0,16]}}'[:16])
    with pytest.raises(ValueError):
        st._load_native(str(bad))


@needs_host
def test_native_token_loader_is_deterministic_and_resumable():
    a = NativeTokenLoader(4, 64, 1024, seed=3, rank=1, nbuf=4, pin_memory=False)
    first = [next(a)["input_ids"].clone() for _ in range(6)]
    assert all(int(x.min()) >= 3 and int(x.max()) < 1024 for x in first) and not torch.equal(first[0], first[1])
    sd = a.state_dict()
    nxt = next(a)["input_ids"].clone()
    b = NativeTokenLoader(4, 64, 1024, seed=3, rank=1, nbuf=4, pin_memory=False)
    assert torch.equal(next(b)["input_ids"], first[0])
    b.load_state_dict(sd)
    assert torch.equal(next(b)["input_ids"], nxt)
    c = NativeTokenLoader(4, 64, 1024, seed=3, rank=2, nbuf=4, pin_memory=False)
    assert not torch.equal(next(c)["input_ids"], first[0])          # ranks see different streams
    a.close(), b.close(), c.close()


def test_python_loader_and_collation():
    ld = SyntheticTokenLoader(2, 16, 100, seed=1, pin_memory=False)
    x = next(ld)
    sd = ld.state_dict()
    y = next(ld)["input_ids"].clone()
    ld.load_state_dict(sd)
    assert torch.equal(next(ld)["input_ids"], y) and x["attention_mask"].all()
    it = iter(FakeTokenizedDataset(8, 50, seed=0))
    feats = [next(it), next(it)]
    feats[1] = {"input_ids": feats[1]["input_ids"][:5], "attention_mask": [1] * 5}
    batch = collate_causal_lm(feats, pad_token_id=2)
    assert batch["input_ids"].shape == (2, 8) and (batch["labels"][1, 5:] == -100).all() and (batch["attention_mask"][1, 5:] == 0).all()
