"""MeshStore: the H5 key layout of extract_mesh_state (Extract_mesh/parse_tfrecord_refactor.py:890-899, 962-965) in a numpy archive,
with the small h5py surface the reference touches — `store[str(i)][key][0]`, `[()]`, `[t0:t1]`, `len(store)`, `create_group(name)`,
`group.create_dataset(key, data=...)`, `close()` — so that Data_Pool (reader) and fvgn_b200.convert (writer) work where h5py is not
installed (this image) and with h5py.File where it is.  One uncompressed .npz per split: `<dir>/<split>.fvgn.npz`, member names
`<trajectory>/<key>`; members are memory-mapped lazily on first access."""
import os
import zipfile

import numpy as np


class _Group(dict):
    """writer side: h5py.Group stand-in"""

    def create_dataset(self, key, data=None, **kw):
        self[key] = np.asarray(data)
        return self[key]


class _LazyGroup:
    """reader side: arrays are loaded from the archive on first access"""

    def __init__(self, store, name, keys):
        self._store, self._name, self._keys, self._cache = store, name, set(keys), {}

    def __getitem__(self, key):
        if key not in self._keys:
            raise KeyError(key)
        a = self._cache.get(key)
        if a is None:
            a = self._cache[key] = self._store._read(f"{self._name}/{key}")
        return a

    def __contains__(self, key):
        return key in self._keys

    def keys(self):
        return sorted(self._keys)

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class MeshStore:
    def __init__(self, path, mode="r"):
        self.path, self.mode = path, mode
        self._groups = {}
        self._zip = None
        if mode == "r":
            self._zip = zipfile.ZipFile(path, "r")
            names = {}
            for n in self._zip.namelist():
                g, _, k = n[:-4].partition("/")  # strip ".npy"
                names.setdefault(g, []).append(k)
            self._groups = {g: _LazyGroup(self, g, ks) for g, ks in names.items()}
        elif mode != "w":
            raise ValueError("mode must be 'r' or 'w'")

    def _read(self, member):
        with self._zip.open(member + ".npy") as f:
            return np.lib.format.read_array(f, allow_pickle=False)

    # h5py.File surface
    def create_group(self, name):
        if self.mode != "w":
            raise IOError("store opened read-only")
        g = self._groups[str(name)] = _Group()
        return g

    def __getitem__(self, name):
        return self._groups[str(name)]

    def __contains__(self, name):
        return str(name) in self._groups

    def __len__(self):
        return len(self._groups)

    def keys(self):
        return sorted(self._groups, key=lambda s: (len(s), s))

    def close(self):
        if self.mode == "w" and self._groups is not None:
            os.makedirs(os.path.dirname(os.path.abspath(self.path)) or ".", exist_ok=True)
            with zipfile.ZipFile(self.path, "w", zipfile.ZIP_STORED, allowZip64=True) as z:
                for g, d in self._groups.items():
                    for k, a in d.items():
                        with z.open(f"{g}/{k}.npy", "w", force_zip64=True) as f:
                            a = np.asarray(a)
                            np.lib.format.write_array(f, a if a.ndim == 0 else np.ascontiguousarray(a), allow_pickle=False)
        if self._zip is not None:
            self._zip.close()
            self._zip = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


def open_store(dataset_dir, split, mode="r"):
    """<dir>/<split>.h5 through h5py when both exist, else <dir>/<split>.fvgn.npz"""
    h5 = os.path.join(dataset_dir, f"{split}.h5")
    if mode == "r" and os.path.exists(h5):
        try:
            import types

            import h5py

            if not isinstance(h5py, types.ModuleType):
                raise ImportError("h5py is not installed")
            return h5py.File(h5, "r")
        except ImportError as e:
            raise RuntimeError(f"{h5} needs h5py, which is not installed: convert to {split}.fvgn.npz with fvgn_b200.convert") from e
    return MeshStore(os.path.join(dataset_dir, f"{split}.fvgn.npz"), mode)
