"""Host-side id shards: same tar/pickle layout as the reference's extractor
(MultiModalLLM/src/tools/extract_image_ids_to_torchdata_parallel.py:76-127)."""
import os
import pickle
import tarfile

import torch

from seed_b200.shards import IdShardWriter, extract_ids_to_shards, image_ids_to_string, read_id_shards


def test_writer_layout_rollover_and_roundtrip(tmp_path):
    keys = iter(f"k{i:04d}" for i in range(100))
    with IdShardWriter(str(tmp_path), rank=3, maxcount=4, key_fn=lambda: next(keys)) as w:
        ids = torch.arange(10 * 32).view(10, 32) % 8192
        texts = [f"caption {i}" for i in range(10)]
        metas = ['{"url": "u%d"}' % i for i in range(5)] + [{"url": "u%d" % i} for i in range(5, 10)]
        w.write_batch(ids, texts, metas)
    part = tmp_path / "part-0003"
    assert sorted(os.listdir(part)) == ["0000000.tar", "0000001.tar", "0000002.tar"]       # 4 + 4 + 2 samples
    with tarfile.open(part / "0000000.tar") as tar:
        members = tar.getmembers()
        assert [m.name for m in members] == ["k0000.pkl", "k0001.pkl", "k0002.pkl", "k0003.pkl"]
        s = pickle.loads(tar.extractfile(members[1]).read())
        assert set(s) == {"image_ids", "text", "metadata"}
        assert s["image_ids"] == list(range(32, 64)) and all(type(v) is int for v in s["image_ids"])
        assert s["text"] == "caption 1" and s["metadata"] == {"url": "u1"}
    got = list(read_id_shards(str(tmp_path)))
    assert len(got) == 10 and [g["__key__"] for g in got] == [f"k{i:04d}" for i in range(10)]
    assert got[7]["metadata"] == {"url": "u7"} and got[9]["image_ids"] == (ids[9] % 8192).tolist()


def test_extract_loop_uses_encode_image_and_one_copy_per_batch(tmp_path):
    class FakeTok:
        device = "cpu"
        calls = 0

        def encode_image(self, image_torch=None):
            self.calls += 1
            return (image_torch.flatten(1).sum(1, keepdim=True).long() + torch.arange(32)[None]) % 8192

    tok = FakeTok()
    batches = [{"pixel_values": torch.ones(3, 3, 4, 4) * (b + 1), "text": ["a", "b", "c"], "metadata": ["{}"] * 3}
               for b in range(2)]
    n = extract_ids_to_shards(tok, batches, str(tmp_path), rank=0)
    assert n == 6 and tok.calls == 2
    got = list(read_id_shards(str(tmp_path)))
    assert got[0]["image_ids"][0] == 48 and got[3]["image_ids"][0] == 96 and len(got[0]["image_ids"]) == 32


def test_wire_string_matches_reference_format():
    s = image_ids_to_string([0, 17, 8191])
    assert s == "<img><img_00000><img_00017><img_08191></img>"
