"""Image-id shards in the reference's on-disk format ("keep the tar/pickle format for
dataset compatibility").

The reference's extractor (MultiModalLLM/src/tools/extract_image_ids_to_torchdata_parallel.py:76-127) writes, per
rank, `save_dir/part-%04d/%07d.tar` through `webdataset.ShardWriter(pattern, maxcount=10000)`; every sample is one
tar member `<uuid4 hex>.pkl` holding `pickle.dumps({'image_ids': [32 ints], 'text': str, 'metadata': dict})`, which
the training reader decodes again (MultiModalLLM/src/data/torchdata_train.py:100-108).  `webdataset` is not a
dependency of this build; the same files are produced with `tarfile` (member names, extension and payload are what
readers key on).  Differences on purpose: ONE device->host copy per batch instead of one `.cpu()` (= one stream
sync) per image, and an optional deterministic key so shards are reproducible.
"""
from __future__ import annotations

import glob
import io
import json
import os
import pickle
import tarfile
import time
import uuid
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import torch


class IdShardWriter:
    def __init__(self, save_dir: str, rank: int = 0, maxcount: int = 10000, key_fn=None):
        self.dir = os.path.join(save_dir, "part-{:04d}".format(rank))
        os.makedirs(self.dir, exist_ok=True)
        self.pattern = os.path.join(self.dir, "%07d.tar")
        self.maxcount = int(maxcount)
        self.key_fn = key_fn or (lambda: uuid.uuid4().hex)
        self.shard, self.count, self.total = 0, 0, 0
        self._tar: Optional[tarfile.TarFile] = None

    # -- webdataset.ShardWriter semantics: a new shard every `maxcount` samples ---------------------------
    def _next_shard(self):
        self.close_shard()
        self._tar = tarfile.open(self.pattern % self.shard, "w")
        self.shard += 1
        self.count = 0

    def close_shard(self):
        if self._tar is not None:
            self._tar.close()
            self._tar = None

    def write_sample(self, image_ids: Sequence[int], text: str, metadata) -> str:
        if self._tar is None or self.count >= self.maxcount:
            self._next_shard()
        if isinstance(metadata, (str, bytes)):
            metadata = json.loads(metadata)            # the reference stores json.loads(metadata) (:124)
        sample = {"image_ids": [int(i) for i in image_ids], "text": text, "metadata": metadata}
        payload = pickle.dumps(sample)
        key = self.key_fn()
        info = tarfile.TarInfo(key + ".pkl")
        info.size = len(payload)
        info.mtime = time.time()
        info.mode = 0o444
        info.uname = info.gname = "bigdata"             # what webdataset's TarWriter stamps
        self._tar.addfile(info, io.BytesIO(payload))
        self.count += 1
        self.total += 1
        return key

    def write_batch(self, image_ids: torch.Tensor, texts: Sequence[str], metadatas: Sequence) -> List[str]:
        """image_ids [B, 32] (device or host) -> B samples; one D2H copy for the whole batch."""
        ids = image_ids.detach().reshape(image_ids.shape[0], -1).to("cpu", torch.int64).tolist()
        if not (len(ids) == len(texts) == len(metadatas)):
            raise ValueError("image_ids, texts and metadatas must have the same length")
        return [self.write_sample(i, t, m) for i, t, m in zip(ids, texts, metadatas)]

    def close(self):
        self.close_shard()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_id_shards(save_dir_or_pattern: str) -> Iterator[Dict]:
    """Yield {'__key__', 'image_ids', 'text', 'metadata'} from shards written by IdShardWriter or by the reference."""
    if os.path.isdir(save_dir_or_pattern):
        files = sorted(glob.glob(os.path.join(save_dir_or_pattern, "part-*", "*.tar")))
    else:
        files = sorted(glob.glob(save_dir_or_pattern))
    for f in files:
        with tarfile.open(f, "r") as tar:
            for m in tar:
                if not m.isfile() or not m.name.endswith(".pkl"):
                    continue
                sample = pickle.loads(tar.extractfile(m).read())
                sample["__key__"] = m.name[:-4]
                yield sample


def extract_ids_to_shards(tokenizer, batches: Iterable[Dict], save_dir: str, rank: int = 0, maxcount: int = 10000,
                          key_fn=None) -> int:
    """The body of the reference's run_worker loop (:104-127) for one rank: every batch dict carries `pixel_values`
    [B,3,224,224], `text` (list of str) and `metadata` (list of json strings or dicts)."""
    with IdShardWriter(save_dir, rank, maxcount, key_fn) as sink:
        with torch.no_grad():
            for batch in batches:
                image_ids = tokenizer.encode_image(image_torch=batch["pixel_values"].to(tokenizer.device))
                sink.write_batch(image_ids, batch["text"], batch["metadata"])
        return sink.total


BOI_TOKEN, EOI_TOKEN, IMG_TOKEN = "<img>", "</img>", "<img_{:05d}>"


def image_ids_to_string(image_ids: Sequence[int]) -> str:
    """the reference's text wire format (torchdata_train.py:21-23,108; scripts/seed_llama_inference_8B.py:98-100)"""
    return BOI_TOKEN + "".join(IMG_TOKEN.format(int(i)) for i in image_ids) + EOI_TOKEN
