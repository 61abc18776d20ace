"""PipelineState on-disk format: the reference's PipelineSerdeManager, byte for byte.

  <base>/save__<alias>/files.json          {"alias": ..., "version": ..., "files": [{"id", "type", "size"}, ...]}
  <base>/save__<alias>/data/<name>.data    the raw bytes of the buffer (little-endian, no header)

(Pipeline/PipelineState/PipelineSerialization.cs: FileDirectory :16-90, FileObject :94-98, BinaryIO :100-183,
PipelineSerdeManager :185-236.)  `type` is the C# element type name (`typeof(T).Name`: float -> "Single", int -> "Int32",
uint -> "UInt32"); `size` is the element count the caller passes.  A tile dumped by a Unity run and one dumped here are
byte-comparable, which is what a future bit-level pin of the oracle needs (tools/dump_chain.py writes one).
"""
import json
import os

import numpy as np

_TYPE_NAMES = {np.dtype(np.float32): "Single", np.dtype(np.int32): "Int32", np.dtype(np.uint32): "UInt32",
               np.dtype(np.float64): "Double", np.dtype(np.uint8): "Byte", np.dtype(np.int16): "Int16",
               np.dtype(np.uint16): "UInt16"}
_INVALID = set('\0/') | {chr(c) for c in range(1, 32)}     # Path.GetInvalidFileNameChars() on Linux/macOS editors ('\0', '/')


def clean_file_name(name):
    """PipelineSerdeManager.CleanFileName, :207-210: split on invalid characters, drop empties, join with '_', trim dots."""
    parts, cur = [], []
    for ch in name:
        if ch in _INVALID:
            if cur:
                parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if cur:
        parts.append("".join(cur))
    return "_".join(parts).rstrip(".")


class PipelineSerdeManager:
    def __init__(self, path, alias, version):
        self.basePath, self.alias, self.version = path, alias, version
        self.fullPath = os.path.join(path, f"save__{alias}", "files.json")
        if os.path.exists(self.fullPath):
            d = json.load(open(self.fullPath))
            self.files = [dict(id=f["id"], type=f["type"], size=int(f["size"])) for f in d.get("files", [])]
            self.alias, self.version = d.get("alias", alias), d.get("version", version)
        else:
            self.files = []
        self._lookup = {f"{f['id']}_{f['type']}": i for i, f in enumerate(self.files)}

    def GetFQN(self, name):
        return os.path.join(self.basePath, f"save__{self.alias}", "data", f"{clean_file_name(name)}.data")

    def _set_count(self, name, type_name, size):
        key = f"{name}_{type_name}"
        if key in self._lookup:
            self.files[self._lookup[key]]["size"] = size
        else:
            self.files.append(dict(id=name, type=type_name, size=size))
            self._lookup[key] = len(self.files) - 1
        os.makedirs(os.path.dirname(self.fullPath), exist_ok=True)
        # JsonUtility.ToJson: compact, fields in declaration order
        with open(self.fullPath, "w") as f:
            f.write(json.dumps({"alias": self.alias, "version": self.version, "files": self.files}, separators=(",", ":")))

    def WriteData(self, data, name, size=None):
        a = np.ascontiguousarray(data)
        if a.dtype not in _TYPE_NAMES:
            raise TypeError(f"unsupported element type {a.dtype}")
        fqn = self.GetFQN(name)
        os.makedirs(os.path.dirname(fqn), exist_ok=True)
        with open(fqn, "wb") as f:
            f.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
        self._set_count(name, _TYPE_NAMES[a.dtype], int(a.size if size is None else size))

    def ReadData(self, name, dtype=np.float32, count=None):
        fqn = self.GetFQN(name)
        if not os.path.exists(fqn):
            return None                                   # "No current file for {name}"
        dt = np.dtype(dtype).newbyteorder("<")
        return np.fromfile(fqn, dtype=dt, count=-1 if count is None else count).astype(np.dtype(dtype))

    def CachedSize(self, name, dtype=np.float32):
        i = self._lookup.get(f"{name}_{_TYPE_NAMES[np.dtype(dtype)]}")
        return -1 if i is None else self.files[i]["size"]
