"""Host-side mirror of the reference's Scala facades over the C ABI (include/dpf.h).

Reference surface kept, same names, argument meaning and error behaviour:
  mclab.deploy.LSHServer                 src/main/scala/mclab/deploy/LSHServer.scala:5-18
  mclab.lsh.LSH(conf)                    src/main/scala/mclab/lsh/LSH.scala:17-166
  mclab.deploy.DensevectorRDFInit        src/main/scala/mclab/deploy/DensevectorRDFInit.scala
  mclab.deploy.SparsevectorRDFInit       src/main/scala/mclab/deploy/SparsevectorRDFInit.scala
  mclab.lsh.vector.Vectors parsers       src/main/scala/mclab/lsh/vector/Vector.scala:162-219, 284-293

The JVM is absent from this image, so this module plays the role of the Scala objects: it owns configuration, hash
function generation/loading (host work in the reference too) and text parsing, and forwards every data-parallel step
to libdpf_b200.so.  Results are Python sets of ids per query (the reference returns Array[Set[AnyRef]] of boxed ints).
Errors: like the reference, querying before fitting prints "need to fit the data first" and yields None.
"""
import os

import numpy as np

from . import _lib as B
from .index import DPFIndex

# ---------------------------------------------------------------------------------------------------------------
# configuration (Typesafe Config subset: `a.b.c = value` lines; TestSettings.scala:6-60 holds the defaults)
# ---------------------------------------------------------------------------------------------------------------
DEFAULT_CONF = """
mclab.confType=lsh
mclab.lsh.name = angle
mclab.lsh.generateByPulling = true
mclab.lsh.IsOrthogonal = true
mclab.lsh.generateMethod = default
mclab.lsh.familyFilePath = "hashFamily/lsh-bestHashFamily-angle-TableNum-10"
mclab.lsh.partitionFamilyFilePath="hashFamily/partition-bestHashFamily-angle"
mclab.lsh.family.pstable.mu = 0.0
mclab.lsh.family.pstable.sigma = 1.0
mclab.lsh.family.pstable.w = 4
mclab.lsh.familySize = 100
mclab.lsh.vectorDim = 100
mclab.lsh.tableNum = 10
mclab.lsh.permutationNum = 3
mclab.lsh.typeOfIndex = original
mclab.lsh.featureDataFormat = sparse
mclab.lshTable.bufferOverflow=500
mclab.lshTable.bucketBits=28
mclab.lshTable.dirNodeSize=32
mclab.lshTable.chainLength = 32
mclab.lsh.partitionBits=3
mclab.insertThreadNum=5
mclab.queryThreadNum=5
mclab.lsh.topK = 10
mclab.lsh.seed = 88387
"""


class Config(dict):
    """`ConfigFactory.parseString(...).withFallback(...)` look-alike."""

    @staticmethod
    def parseString(text):
        c = Config()
        for line in text.splitlines():
            line = line.split("#")[0].strip().lstrip("|").strip()
            if not line or "=" not in line:
                continue
            k, v = line.split("=", 1)
            c[k.strip()] = v.strip().strip('"')
        return c

    def withFallback(self, other):
        out = Config(other)
        out.update(self)
        return out

    def getString(self, k):
        return str(self[k])

    def getInt(self, k):
        return int(self[k])

    def getDouble(self, k):
        return float(self[k])

    def getBoolean(self, k):
        return str(self[k]).lower() == "true"


testBaseConf = Config.parseString(DEFAULT_CONF)

_TRANSFORMS = {"original": B.KEY_ORIGINAL, "sampling": B.KEY_SAMPLING, "continueBitsCount": B.KEY_CONTINUE_BITS,
               "angleNewMethod": B.KEY_ANGLE_NEW}


# ---------------------------------------------------------------------------------------------------------------
# vectors and text parsers (Vector.scala)
# ---------------------------------------------------------------------------------------------------------------
class DenseVector:
    def __init__(self, vectorId, values):
        self.vectorId, self.values = int(vectorId), np.asarray(values, np.float64)

    @property
    def size(self):
        return len(self.values)

    def toArray(self):
        return self.values

    def __repr__(self):
        return "[" + ",".join(repr(float(v)) for v in self.values) + "]"


class SparseVector:
    def __init__(self, vectorId, size, indices, values):
        self.vectorId, self.size = int(vectorId), int(size)
        self.indices, self.values = np.asarray(indices, np.int32), np.asarray(values, np.float64)
        if len(self.indices) != len(self.values):
            raise ValueError(f"indices length: {len(self.indices)}, values length: {len(self.values)}")

    def toArray(self):
        out = np.zeros(self.size)
        out[self.indices] = self.values
        return out

    def __repr__(self):
        return "(%d,%d,[%s],[%s])" % (self.vectorId, self.size, ",".join(str(int(i)) for i in self.indices),
                                      ",".join(repr(float(v)) for v in self.values))


class Vectors:
    @staticmethod
    def parseDense(line):
        """`[1,[0.1,0.2,0.4,0.9]]` -> (id, values) (Vector.scala:215-219)."""
        parts = line.replace(" ", "").replace("[", "").replace("]", "").split(",")
        return int(parts[0]), np.array([float(v) for v in parts[1:]], np.float64)

    @staticmethod
    def _three(parts_text, open_ch, close2):
        parts = parts_text.split(",[")
        if len(parts) != 3:
            raise ValueError(f"cannot parse {parts_text}")
        vid, size = (int(v) for v in parts[0].replace(open_ch, "").split(","))
        idx = [int(v) for v in parts[1].replace("]", "").split(",") if v != ""]
        val = [float(v) for v in parts[2].replace(close2, "").split(",") if v != ""]
        return vid, size, np.array(idx, np.int32), np.array(val, np.float64)

    @staticmethod
    def fromString(s):
        """`(3,3,[0,1,2],[1.0,2.0,3.0])` (Vector.scala:162-175)."""
        return Vectors._three(s.strip(), "(", "])")

    @staticmethod
    def fromPythonString(s):
        """`[1, 3, [1, 2, 3], [1.0, 2.0, 3.0]]` (Vector.scala:194-208)."""
        return Vectors._three(s.strip().replace(" ", ""), "[", "]]")

    @staticmethod
    def fromStringDense(s):
        return np.array([float(v) for v in s.split(",")], np.float64)

    @staticmethod
    def analysisKNN(line, k):
        """top-k neighbour ids `[1,30,19,...]` (Vector.scala:284-293)."""
        parts = line.replace(" ", "").split(",")
        if k > len(parts):
            raise ValueError(f"cannot parse {line}")
        return np.array([int(p.replace("[", "").replace("]", "")) for p in parts[:k]], np.int32)


def _first_line_dim(path):
    with open(path) as f:
        for line in f:
            if line.strip():
                return len(Vectors.parseDense(line)[1])
    return 0


def load_dense_file(path, d=None):
    """Whole-file reader of the `[id,[v...]]` format through the library's native parser (dpf_parse_dense_file):
    rows in file order, ids ignored (quirk Q14).  d defaults to the width of the first line."""
    import ctypes as C
    lib = B.load()
    d = d or _first_line_dim(path)
    if d == 0:
        return np.zeros((0, 0))
    n = C.c_int64(0)
    rc = lib.dpf_parse_dense_file(str(path).encode(), d, None, 0, C.byref(n))
    if rc != B.OK:
        raise ValueError(f"cannot parse {path}: {lib.dpf_strerror(rc).decode()}")
    X = np.empty((n.value, d), np.float64)
    rc = lib.dpf_parse_dense_file(str(path).encode(), d, C.c_void_p(X.ctypes.data), n.value, C.byref(n))
    if rc != B.OK:
        raise ValueError(f"cannot parse {path}: {lib.dpf_strerror(rc).decode()}")
    return X


def load_sparse_file(path):
    """`[id, size, [idx], [val]]` rows -> (CSR (indptr int64, indices int32, values f64), dim) through
    dpf_parse_sparse_file; indices ascending per row (BitSet iteration order, SimilarityCalculator.scala:19-25)."""
    import ctypes as C
    lib = B.load()
    n, nnz, dim = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    rc = lib.dpf_parse_sparse_file(str(path).encode(), None, None, None, 0, 0, C.byref(n), C.byref(nnz), C.byref(dim))
    if rc != B.OK:
        raise ValueError(f"cannot parse {path}: {lib.dpf_strerror(rc).decode()}")
    indptr = np.zeros(n.value + 1, np.int64)
    idx = np.empty(nnz.value, np.int32)
    val = np.empty(nnz.value, np.float64)
    rc = lib.dpf_parse_sparse_file(str(path).encode(), C.c_void_p(indptr.ctypes.data), C.c_void_p(idx.ctypes.data),
                                   C.c_void_p(val.ctypes.data), n.value, max(nnz.value, 1), C.byref(n), C.byref(nnz),
                                   C.byref(dim))
    if rc != B.OK:
        raise ValueError(f"cannot parse {path}: {lib.dpf_strerror(rc).decode()}")
    return (indptr, idx, val), dim.value


# ---------------------------------------------------------------------------------------------------------------
# LSH: owns the hash functions (tableIndexGenerators) — generated or loaded on the host, as in the reference
# ---------------------------------------------------------------------------------------------------------------
class LSH:
    """`new LSH(conf)` (LSH.scala:17-82).  Functions are generated with a seeded numpy RNG (`mclab.lsh.seed`; the
    reference's RNG is unseeded, quirk Q7) or loaded from `familyFilePath` (`generateMethod = fromfile`,
    AngleHashFamily.scala:158-177, which yields lines/chainLength chains and ignores permutationNum, quirk Q8)."""

    def __init__(self, conf):
        self.conf = conf
        self.name = conf.getString("mclab.lsh.name")
        self.typeOfIndex = conf.getString("mclab.lsh.typeOfIndex")
        d = conf.getInt("mclab.lsh.vectorDim")
        k = conf.getInt("mclab.lshTable.chainLength")
        tn = conf.getInt("mclab.lsh.tableNum")
        pn = conf.getInt("mclab.lsh.permutationNum")
        fs = conf.getInt("mclab.lsh.familySize")
        seed = int(conf.get("mclab.lsh.seed", 88387))
        self.b = self.w = None
        if conf.getString("mclab.lsh.generateMethod") == "fromfile":
            key = "mclab.lsh.partitionFamilyFilePath" if conf.getString("mclab.confType") == "partition" \
                else "mclab.lsh.familyFilePath"
            rows, bs, ws = [], [], []
            for line in open(conf.getString(key)):
                if not line.strip():
                    continue
                if self.name == "pStable":
                    vs, b, w = line.strip().split(";")
                    bs.append(float(b)); ws.append(int(w))
                else:
                    vs = line.strip()
                _, size, idx, val = Vectors.fromString(vs)
                row = np.zeros(size); row[idx] = val
                rows.append(row)
            A = np.stack(rows)
            nchains = len(rows) // k
            self.A, self.chain = A[:nchains * k], np.arange(nchains * k, dtype=np.int32).reshape(nchains, k)
            if self.name == "pStable":
                self.b, self.w = np.array(bs[:nchains * k]), np.array(ws[:nchains * k], np.int32)
        elif self.name == "angle":
            from . import synth
            self.A, self.chain = synth.angle_family(d, fs, tn, pn, k, seed, conf.getBoolean("mclab.lsh.generateByPulling"))
        elif self.name == "pStable":
            rng = np.random.default_rng(seed)
            mu, sigma = conf.getDouble("mclab.lsh.family.pstable.mu"), conf.getDouble("mclab.lsh.family.pstable.sigma")
            w = conf.getInt("mclab.lsh.family.pstable.w")
            self.A = mu + sigma * rng.standard_normal((fs, d))                  # PStableHashFamily.scala:37-57
            self.b = rng.random(fs) * w
            self.w = np.full(fs, w, np.int32)
            self.chain = rng.integers(0, fs, (tn, k)).astype(np.int32)           # pick: no permutations (:66-78)
        else:
            raise ValueError(f"{self.name} is not a valid family name")
        self.tableIndexGenerators = [self.chain[t] for t in range(self.chain.shape[0])]

    @property
    def family_kind(self):
        return B.FAMILY_PSTABLE if self.name == "pStable" else B.FAMILY_ANGLE

    def calculateIndex(self, vector, tableId=-1):
        """LSH.calculateIndex (LSH.scala:93-166): keys of all tables (tableId < 0) or `Array(key)` of one table, with the
        configured typeOfIndex transform.  Evaluated by the library (dpf_hash_dense / dpf_hash_csr) on a private handle
        that holds only the hash functions."""
        if getattr(self, "_hasher", None) is None:
            d = self.conf.getInt("mclab.lsh.vectorDim")
            self._hasher = DPFIndex(d=d, L=self.chain.shape[0], k=self.chain.shape[1], pb=0, family_kind=self.family_kind,
                                    key_transform=_TRANSFORMS[self.typeOfIndex],
                                    device=int(self.conf.get("mclab.gpu.device", 0)))
            self._hasher.set_family(self.A, self.chain, self.b, self.w)
        if isinstance(vector, SparseVector):
            keys, _ = self._hasher.hash_csr(np.array([0, len(vector.indices)], np.int64), vector.indices, vector.values)
        else:
            values = vector.values if isinstance(vector, DenseVector) else np.asarray(vector, np.float64)
            keys, _ = self._hasher.hash_dense(values[None, :])
        keys = keys[:, 0]
        return keys.copy() if tableId < 0 else keys[tableId:tableId + 1].copy()


class LSHServer:
    """Global engine holder (LSHServer.scala:5-18)."""
    lshEngine = None
    isUseDense = False

    @staticmethod
    def getLSHEngine():
        return LSHServer.lshEngine

    @staticmethod
    def getisUseDense():
        return LSHServer.isUseDense


# ---------------------------------------------------------------------------------------------------------------
# facades
# ---------------------------------------------------------------------------------------------------------------
class _RDFInit:
    """Shared body of the dense and sparse singleton objects."""
    _dense = True

    def __init__(self):
        self.index = None
        self.conf = None
        self.tableNum = self.permutationNum = 0
        self.vectors = None

    # DensevectorRDFInit.initializeRDFHashMap (DensevectorRDFInit.scala:50-118)
    def initializeRDFHashMap(self, conf):
        if LSHServer.lshEngine is None:
            LSHServer.lshEngine = LSH(conf)
        lsh = LSHServer.lshEngine
        self.conf = conf
        self.tableNum, self.permutationNum = conf.getInt("mclab.lsh.tableNum"), conf.getInt("mclab.lsh.permutationNum")
        pb = conf.getInt("mclab.lsh.partitionBits")
        L = lsh.chain.shape[0]
        from . import synth
        # one private LocalitySensitivePartitioner per table (confForPartitioner: vectorDim=32, chainLength=partitionBits)
        pseed = int(conf.get("mclab.lsh.seed", 88387)) + 1
        if lsh.name == "pStable":    # confForPartitioner falls back to the main conf: the partitioner chains are pStable too
            Ap, pb_b, pb_w = synth.pstable_partitioner_family(
                L, pb, conf.getDouble("mclab.lsh.family.pstable.mu"), conf.getDouble("mclab.lsh.family.pstable.sigma"),
                conf.getInt("mclab.lsh.family.pstable.w"), pseed)
        else:
            Ap, pb_b, pb_w = synth.partitioner_family(L, pb, pseed), None, None
        self.close()
        self.index = DPFIndex(d=conf.getInt("mclab.lsh.vectorDim"), L=L, k=lsh.chain.shape[1], pb=pb,
                              bucket_bits=conf.getInt("mclab.lshTable.bucketBits"),
                              dir_node_size=conf.getInt("mclab.lshTable.dirNodeSize"),
                              bucket_overflow=conf.getInt("mclab.lshTable.bufferOverflow"),
                              family_kind=lsh.family_kind, key_transform=_TRANSFORMS[lsh.typeOfIndex],
                              device=int(conf.get("mclab.gpu.device", 0)), rank=int(conf.get("mclab.gpu.rank", 0)),
                              world=int(conf.get("mclab.gpu.world", 1)))
        self.index.set_family(lsh.A, lsh.chain, lsh.b, lsh.w)
        self.index.set_partitioners(Ap, pb_b, pb_w)
        self.partitioners = Ap

    @property
    def vectorIdToVector(self):
        """The reference's dataTable (id -> vector), here the list of vectors in id order (get = indexing, size = len)."""
        return self.vectors

    @property
    def vectorDatabase(self):
        """The reference's `Array[RandomDrawTreeMap]`, one per table: here per-table views of the flat forest
        (`dump()` = canonical leaf buckets of that table)."""
        if self.index is None:
            return []
        ix = self.index
        return [type("TableView", (), {"tableId": t, "dump": (lambda self_, t=t: ix.dump_buckets(t)),
                                       "size": (lambda self_: len(ix))})() for t in range(ix.L)]

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None

    def clearAndClose(self):
        self.close()
        self.vectors = None

    def _sets(self, off, ids):
        return [set(ids[off[i]:off[i + 1]].tolist()) for i in range(len(off) - 1)]

    def getDtAndHtNumDistribution(self):
        """(dataTable distribution, hashTable distribution averaged over tables) (DensevectorRDFInit.scala:515-530)."""
        n = len(self.index)
        return np.array([float(n)]), self.index.stats()["occupancy"]

    @staticmethod
    def getTopKGroundTruth(filename, K):
        return [set(Vectors.analysisKNN(line, K).tolist()) for line in open(filename) if line.strip()]


class _DensevectorRDFInit(_RDFInit):
    def _fit(self, fileName, conf):
        LSHServer.isUseDense = True
        self.initializeRDFHashMap(conf)
        X = load_dense_file(fileName)
        self.index.fit_dense(X)
        self.vectors = [DenseVector(i, X[i]) for i in range(len(X))]
        return self.vectors

    # newFastFit / newMultiThreadFit (DensevectorRDFInit.scala:127-206): the thread count is irrelevant on the GPU;
    # both give the forest of sequential ascending-id insertion
    def newFastFit(self, fileName, conf):
        return self._fit(fileName, conf)

    def newMultiThreadFit(self, fileName, conf):
        return self._fit(fileName, conf)

    def fitArray(self, X, conf):
        """Same as newMultiThreadFit for vectors already in memory (n x d array)."""
        LSHServer.isUseDense = True
        self.initializeRDFHashMap(conf)
        self.index.fit_dense(X)

    def _stack(self, vecs):
        return np.stack([v.values if isinstance(v, DenseVector) else np.asarray(v, np.float64) for v in vecs])

    def querySingleKey(self, queryKey, denseVector, steps=0, L=None):
        r = self.queryBatch([queryKey], [denseVector], steps, L)
        return None if r is None else r[0]

    def queryBatch(self, queryArray, denseVectorArray, steps=0, L=None):
        return self.NewMultiThreadQueryBatch(queryArray, denseVectorArray, steps)

    def NewMultiThreadQueryBatch(self, queryArray, denseVectorArray=None, steps=0, queryThreadNum=5):
        """(ids, vectors, steps, threads) -> candidate sets (DensevectorRDFInit.scala:335-360); the overload that takes
        only vectors first inserts them with fresh ids and then queries (:372-399)."""
        if self.index is None or len(self.index) == 0:
            print("need to fit the data first")             # DensevectorRDFInit.scala:422-424
            return None
        if denseVectorArray is None or isinstance(denseVectorArray, int):
            if isinstance(denseVectorArray, int):
                steps = denseVectorArray
            vecs = self._stack(queryArray)
            first = len(self.index)
            self.index.fit_dense(vecs)                       # newMultiThreadFit(querySVArray, threadNum) (:215-251)
            qids = np.arange(first, first + len(vecs), dtype=np.int32)
            return self._sets(*self.index.query_candidates_dense(vecs, qids, steps))
        Q = self._stack(denseVectorArray)
        return self._sets(*self.index.query_candidates_dense(Q, np.asarray(queryArray, np.int32), steps))

    def query(self, queryKeyArray, querySVArray, steps=0, queryThreadNum=10, queryThreadPool=None):
        return self.NewMultiThreadQueryBatch(queryKeyArray, querySVArray, steps)

    def topKAndPrecisionScore(self, allDenseVectors, groundTruth, conf, steps=0, queryThreadPool=None, metric=B.METRIC_DOT):
        """(topK ids per query, precision) — re-rank by descending dot product (DensevectorRDFInit.scala:472-507)."""
        nq, K = len(groundTruth), conf.getInt("mclab.lsh.topK")
        Q = self._stack(allDenseVectors[:nq])
        ids, _ = self.index.query_topk_dense(Q, np.arange(nq, dtype=np.int32), steps, K, metric)
        score = 0.0
        out = []
        for i in range(nq):
            got = [int(v) for v in ids[i] if v >= 0]
            out.append(got)
            score += sum(1 for v in got if v in groundTruth[i]) / nq
        return out, score / K


class _SparsevectorRDFInit(_RDFInit):
    _dense = False

    def _fit(self, fileName, conf):
        LSHServer.isUseDense = False
        self.initializeRDFHashMap(conf)
        (indptr, idx, val), _ = load_sparse_file(fileName)
        self.index.fit_csr(indptr, idx, val)
        self.csr = (indptr, idx, val)
        return [val[indptr[i]:indptr[i + 1]] for i in range(len(indptr) - 1)]   # Array[Array[Double]] of the values

    def newFastFit(self, fileName, conf):
        return self._fit(fileName, conf)

    def newMultiThreadFit(self, fileName, conf):
        return self._fit(fileName, conf)

    def fitCSR(self, indptr, indices, values, conf):
        LSHServer.isUseDense = False
        self.initializeRDFHashMap(conf)
        self.index.fit_csr(indptr, indices, values)

    def NewMultiThreadQueryBatch(self, queryArray, steps=0, queryThreadNum=5):
        """(ids, steps, threads): id-based, probe-less (SparsevectorRDFInit.scala:324-348, :407); a list of SparseVector
        is first inserted, then queried by id (:360-387)."""
        if self.index is None or len(self.index) == 0:
            print("need to fit the data first")
            return None
        if len(queryArray) and isinstance(queryArray[0], SparseVector):
            indptr = np.zeros(len(queryArray) + 1, np.int64)
            for i, v in enumerate(queryArray):
                indptr[i + 1] = indptr[i] + len(v.indices)
            idx = np.concatenate([v.indices for v in queryArray])
            val = np.concatenate([v.values for v in queryArray])
            first = len(self.index)
            self.index.fit_csr(indptr, idx, val)
            qids = np.arange(first, first + len(queryArray), dtype=np.int32)
        else:
            qids = np.asarray(queryArray, np.int32)
        return self._sets(*self.index.query_candidates_by_id(qids, steps))

    def query(self, queryKeyArray, querySVArray, steps=0, queryThreadNum=10):
        """Vector-based sparse query: steps but no probes (RandomDrawTreeMap.java:686-732)."""
        indptr = np.zeros(len(querySVArray) + 1, np.int64)
        for i, v in enumerate(querySVArray):
            indptr[i + 1] = indptr[i] + len(v.indices)
        idx = np.concatenate([v.indices for v in querySVArray])
        val = np.concatenate([v.values for v in querySVArray])
        return self._sets(*self.index.query_candidates_csr(indptr, idx, val, np.asarray(queryKeyArray, np.int32), steps))


DensevectorRDFInit = _DensevectorRDFInit()
SparsevectorRDFInit = _SparsevectorRDFInit()
