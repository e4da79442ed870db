"""Synthetic Derm7pt-shaped dataset so the reference's training scripts can run UNCHANGED without the real data.

The scripts build their dataset as ``datasets.__dict__[args.data_name](args, **kwargs)`` (src/utils/misc.py:433); the
import hook (``_hook.py``) registers this class in ``src.utils.data.datasets`` under the name ``SM3SyntheticPairs``, so
``--data-name SM3SyntheticPairs --data-path none`` selects it.  It follows ``SevenPCBaseDataset``
(src/utils/data/datasets.py:477-535): ``__getitem__`` -> ``(derm, clinic, label[8])`` with ``data_trans`` applied to a
PIL image per modality (a list of views under NViewsTransform), optionally preceded by the sample index
(``return_index=True``, tools/mlc_train.py:328-330).  Images are deterministic per index: the clinic image is a
perturbed copy of the derm image, so cross-modal positives are learnable.  The real class cannot be constructed under
numpy 2.x (``np.alltrue``, datasets.py:143) and the Derm7pt files are not available offline.

Length: env ``SM3_SYNTH_LEN`` (default 413, the Derm7pt training split size); image side: ``SM3_SYNTH_SIDE`` (default 256).
"""
import os

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

NUM_CLASSES = (5, 3, 2, 3, 3, 3, 3, 2)


class SM3SyntheticPairs(Dataset):
    LABEL_ORD = ["DIAG", "PN", "BWV", "VS", "PIG", "STR", "DaG", "RS"]

    def __init__(self, args, data_trans=None, mode="train", return_index=False, return_img_path=False):
        super().__init__()
        self.data_trans = data_trans
        self.mode = mode
        self.return_index = return_index
        self.return_img_path = return_img_path
        self.length = int(os.environ.get("SM3_SYNTH_LEN", "413"))
        self.side = int(os.environ.get("SM3_SYNTH_SIDE", "256"))
        self.seed = int(getattr(args, "seed", 3407)) + {"train": 0, "val": 1, "test": 2}.get(mode, 3)

    def __len__(self):
        return self.length

    def _images(self, index):
        rng = np.random.default_rng(self.seed * 1000003 + index)
        low = rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)          # smooth blobs, not white noise
        derm = np.asarray(Image.fromarray(low).resize((self.side, self.side), Image.BILINEAR), dtype=np.int16)
        clinic = np.clip(derm + rng.integers(-40, 41, size=derm.shape, dtype=np.int16), 0, 255)
        return Image.fromarray(derm.astype(np.uint8)), Image.fromarray(clinic.astype(np.uint8))

    def _apply(self, img):
        if self.data_trans is None:
            return img
        if isinstance(self.data_trans, list):
            return [t(img) for t in self.data_trans]
        return self.data_trans(img)

    def __getitem__(self, index):
        derm, clinic = self._images(index)
        rng = np.random.default_rng(self.seed * 7919 + index)
        label = torch.as_tensor([int(rng.integers(0, c)) for c in NUM_CLASSES])
        sample = (self._apply(derm), self._apply(clinic), label)
        if self.return_index:
            return index, sample
        if self.return_img_path:
            return (f"synthetic://derm/{index}", f"synthetic://clinic/{index}"), sample
        return sample
