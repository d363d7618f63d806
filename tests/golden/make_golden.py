"""Generates the committed golden fixtures from the reference's own test resources.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (committed):
  angle_family_tablenum10.npz   — the 10 x 32 pinned angle functions (100-d) of
                                  src/test/resources/hashFamily/lsh-bestHashFamily-angle-TableNum-10, with the
                                  function ids, so draw-with-replacement structure is preserved
  best_family_angle.npz         — src/test/resources/hashFamily/bestHashFamily-angle (320 x 100-d)
  partition_family_angle.npz    — partition-bestHashFamily-angle-TableNum-1 (2 x 32-d) and
                                  theBestHashFamilyForPartition-angle (62 x 100-d)
The text format is the SparseVector.toString form "(id,size,[indices],[values])" parsed as
Vector.scala:162-175 (Vectors.fromString) does.
"""
import os

import numpy as np

REF = "/root/reference/src/test/resources/hashFamily"
OUT = os.path.dirname(os.path.abspath(__file__))


def from_string(line):
    """Vectors.fromString (Vector.scala:162-175)."""
    parts = line.strip().split(",[")
    assert len(parts) == 3, line[:80]
    vid, size = (int(v) for v in parts[0].replace("(", "").split(","))
    idx = [int(v) for v in parts[1].replace("]", "").split(",") if v != ""]
    val = [float(v) for v in parts[2].replace("])", "").split(",") if v != ""]
    return vid, size, np.array(idx, np.int32), np.array(val, np.float64)


def load(path):
    ids, rows = [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            try:
                vid, size, idx, val = from_string(line)
            except AssertionError:
                # theBestHashFamilyForPartition-angle line 32 is two records run together; Vectors.fromString
                # throws on it in the reference too ("cannot parse"), so it is skipped here
                print("skipping unparsable line in", os.path.basename(path))
                continue
            dense = np.zeros(size, np.float64)
            dense[idx] = val
            ids.append(vid)
            rows.append(dense)
    return np.array(ids, np.int32), np.stack(rows)


if __name__ == "__main__":
    ids, rows = load(os.path.join(REF, "lsh-bestHashFamily-angle-TableNum-10"))
    assert rows.shape == (320, 100)
    np.savez_compressed(os.path.join(OUT, "angle_family_tablenum10.npz"), ids=ids, rows=rows)
    ids2, rows2 = load(os.path.join(REF, "bestHashFamily-angle"))
    np.savez_compressed(os.path.join(OUT, "best_family_angle.npz"), ids=ids2, rows=rows2)
    pid1, prow1 = load(os.path.join(REF, "partition-bestHashFamily-angle-TableNum-1"))
    pid2, prow2 = load(os.path.join(REF, "theBestHashFamilyForPartition-angle"))
    np.savez_compressed(os.path.join(OUT, "partition_family_angle.npz"), ids32=pid1, rows32=prow1, ids100=pid2,
                        rows100=prow2)
    print("family:", rows.shape, "distinct ids", len(set(ids.tolist())), "| best:", rows2.shape,
          len(set(ids2.tolist())), "| partition:", prow1.shape, prow2.shape)
