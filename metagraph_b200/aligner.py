"""Host-side mirror of the reference's aligner interface for the `metagraph align` hot path.

  BOSSTable          ~ boss::BOSS construction result (W / last / F)   [index build, untimed]
  DBGSuccinctIndex   ~ DBGSuccinct resident in HBM (mgb_index_t)
  B200Aligner        ~ IDBGAligner (graph/alignment/dbg_aligner.hpp:20-39): get_graph(),
                       get_config(), align_batch(batch, callback), align(query)
  AlignmentResults / Alignment ~ graph/alignment/alignment.hpp:132-406
  format_alignment   ~ cli/align.cpp:254-307 (TSV branch)

Everything computes through the C-ABI in include/mgb.h (metagraph_b200/_lib/libmgb.so).
"""
import ctypes

import numpy as np

from . import _lib
from .config import DBGAlignerConfig, mgb_config_t

_OPS = "SX=DIG"


def _pack(seqs):
    """list of str/bytes -> (uint8 buffer, uint64 offsets)"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, np.uint8)
    return np.ascontiguousarray(buf), offsets


class BOSSTable:
    """W / last / F arrays of a BOSS table (position 0 is the placeholder)."""

    def __init__(self, k, W, last, F, alphabet=0):
        self.k = int(k)
        self.W = np.ascontiguousarray(W, dtype=np.uint8)
        self.last = np.ascontiguousarray(last, dtype=np.uint8)
        self.F = np.ascontiguousarray(F, dtype=np.uint64)
        self.alphabet = alphabet

    @property
    def num_edges(self):
        return len(self.W) - 1

    @classmethod
    def from_sequences(cls, k, seqs, threads=0, force_source_dummies=False, lib=None, packed=None,
                       alphabet=0):
        """alphabet: 0 = DNA ("$ACGT"), 1 = protein ("$ABCDEFGHIJKLMNOPQRSTUVWYZX")"""
        L = _lib.load_library(lib)
        buf, offsets = packed if packed is not None else _pack(seqs)
        b = _lib.mgb_boss_t()
        rc = L.mgb_boss_build(buf.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k, int(alphabet),
                              int(force_source_dummies), threads, ctypes.byref(b))
        if rc:
            raise _lib.MgbError(rc, "mgb_boss_build: " + L.mgb_boss_last_error().decode())
        try:
            n1 = b.n_plus_1
            W = np.ctypeslib.as_array(b.W, shape=(n1,)).copy()
            last = np.ctypeslib.as_array(b.last, shape=(n1,)).copy()
            F = np.array([b.F[i] for i in range(27 if alphabet == 1 else 5)], dtype=np.uint64)
        finally:
            L.mgb_boss_free(ctypes.byref(b))
        return cls(k, W, last, F, alphabet=int(alphabet))

    @classmethod
    def from_dbg(cls, path, lib=None):
        """Loads a graph written by the reference (`metagraph build` -> *.dbg; DBGSuccinct::load,
        dbg_succinct.cpp:690-712). Returns the table; .mode / .state carry the file's graph mode and BOSS state."""
        L = _lib.load_library(lib)
        b = _lib.mgb_boss_t()
        mode, state = ctypes.c_int(-1), ctypes.c_int(-1)
        rc = L.mgb_dbg_load(str(path).encode(), ctypes.byref(b), ctypes.byref(mode), ctypes.byref(state))
        if rc:
            raise _lib.MgbError(rc, "mgb_dbg_load: " + L.mgb_dbg_last_error().decode())
        try:
            n1 = b.n_plus_1
            W = np.ctypeslib.as_array(b.W, shape=(n1,)).copy()
            last = np.ctypeslib.as_array(b.last, shape=(n1,)).copy()
            F = np.array([b.F[i] for i in range(27 if b.alphabet == 1 else 5)], dtype=np.uint64)
            t = cls(b.k, W, last, F, alphabet=int(b.alphabet))
        finally:
            L.mgb_boss_free(ctypes.byref(b))
        t.mode, t.state = mode.value, state.value
        return t

    def dummy_mask(self, lib=None):
        """valid-edge bytes as DBGSuccinct::mask_dummy_kmers computes them"""
        L = _lib.load_library(lib)
        b = _lib.mgb_boss_t()
        b.n_plus_1 = len(self.W)
        b.W = self.W.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        b.last = self.last.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        for i in range(len(self.F)):
            b.F[i] = int(self.F[i])
        b.k = self.k
        b.alphabet = self.alphabet
        valid = np.zeros(len(self.W), np.uint8)
        _lib.check(L, L.mgb_boss_mask_dummy(ctypes.byref(b), valid.ctypes.data))
        return valid


class DBGSuccinctIndex:
    """BOSS index resident in one GPU's HBM (DBGSuccinct's role for the aligner)."""

    def __init__(self, boss, valid=None, device=0, suffix_len=0, lib=None, mode=None):
        self._L = _lib.load_library(lib)
        self.k = boss.k
        if mode is None:        # a table read from a .dbg file carries the file's mode
            mode = getattr(boss, "mode", 0)
        h = ctypes.c_void_p()
        v = None
        if valid is not None:
            v = np.ascontiguousarray(valid, dtype=np.uint8)
        _lib.check(self._L, self._L.mgb_index_create(
            boss.W.ctypes.data, boss.last.ctypes.data, len(boss.W), boss.F.ctypes.data,
            v.ctypes.data if v is not None else None, boss.k, boss.alphabet, suffix_len, device,
            ctypes.byref(h)))
        self.handle = h
        if mode:     # DeBruijnGraph::Mode: 1 = CANONICAL (graph built from the sequences and their reverse complements)
            try:
                _lib.check(self._L, self._L.mgb_index_set_mode(h, int(mode)))
            except Exception:
                self.close()
                raise
        self.mode = mode

    def close(self):
        if self.handle:
            self._L.mgb_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_k(self):
        return self.k

    @property
    def num_edges(self):
        return self._L.mgb_index_num_edges(self.handle)

    @property
    def device_bytes(self):
        return self._L.mgb_index_device_bytes(self.handle)

    def map_to_nodes_sequentially(self, seqs, packed=None):
        """node ids of all k-mers of each sequence (0 = missing); list of uint64 arrays"""
        buf, offsets = packed if packed is not None else _pack(seqs)
        lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
        nk = np.maximum(lens - self.k + 1, 0)
        out = np.zeros(int(nk.sum()) + 1, dtype=np.uint64)
        _lib.check(self._L, self._L.mgb_map_to_nodes(self.handle, buf.ctypes.data, offsets.ctypes.data,
                                                     len(lens), out.ctypes.data))
        res, p = [], 0
        for n in nk:
            res.append(out[p:p + int(n)])
            p += int(n)
        return res


class Alignment:
    __slots__ = ("orientation", "score", "offset", "query_begin", "query_len", "nodes", "sequence", "cigar")

    def get_cigar_string(self):
        return "".join("%d%s" % (int(x) >> 3, _OPS[int(x) & 7]) for x in self.cigar)

    def get_num_matches(self):
        return int(sum(int(x) >> 3 for x in self.cigar if int(x) & 7 == 2))

    def __repr__(self):   # fmt::formatter<Alignment> (alignment.hpp:418-435)
        return "%s\t%s\t%d\t%d\t%s\t%d" % ("-" if self.orientation else "+", self.sequence, self.score,
                                           self.get_num_matches(), self.get_cigar_string(), self.offset)


_COMP = {}
for _f, _t in zip("ABCDGHKMRTUVY", "TVGHCDMKYAABR"):      # COMPL_TAB (seq_tools/reverse_complement.hpp:31-48)
    _COMP[_f] = _t
    _COMP[_f.lower()] = _t.lower()
_COMP["`"] = "@"


def reverse_complement(seq):
    return "".join(_COMP.get(c, c) for c in reversed(seq))


def _json_value(v):
    """Json::StreamWriterBuilder with indentation "" (jsoncpp): sorted keys, no blanks, doubles as %.17g"""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        t = "%.17g" % v
        return t if any(c in t for c in ".en") else t + ".0"
    if isinstance(v, str):
        out = ['"']
        for ch in v:
            if ch in '"\\':
                out.append("\\" + ch)
            elif ch == "\n":
                out.append("\\n")
            elif ch == "\t":
                out.append("\\t")
            elif ord(ch) < 0x20:
                out.append("\\u%04x" % ord(ch))
            else:
                out.append(ch)
        out.append('"')
        return "".join(out)
    if isinstance(v, list):
        return "[" + ",".join(_json_value(x) for x in v) + "]"
    return "{" + ",".join(_json_value(k) + ":" + _json_value(v[k]) for k in sorted(v)) + "}"


def _path_json(nodes, cigar, node_size, query_view, offset, label):
    """path_json (alignment.cpp:704-880), derived there from GraphAligner's vg.proto"""
    ops = [(int(x) & 7, int(x) >> 3) for x in cigar]          # (Cigar::Operator, length); 0 = CLIPPED
    it = 1 if ops and ops[0][0] == 0 else 0
    cigar_offset = 0
    qpos = 0
    path = {"mapping": []}
    cur_pos = offset
    position = {"node_id": int(nodes[0])}
    if cur_pos:
        position["offset"] = cur_pos
    mapping = {"position": position, "edit": []}
    while cur_pos < node_size and it < len(ops):                # the first node
        op, n = ops[it]
        if op == 0:                                             # trailing clipping
            it += 1
            cigar_offset = 0
            continue
        next_pos = min(node_size, cur_pos + (n - cigar_offset))
        next_size = next_pos - cur_pos
        edit = {}
        if op == 1:                                             # MISMATCH
            edit = {"from_length": next_size, "to_length": next_size, "sequence": query_view[qpos:qpos + next_size]}
            qpos += next_size
        elif op == 4:                                           # INSERTION
            edit = {"to_length": next_size, "sequence": query_view[qpos:qpos + next_size]}
            qpos += next_size
            next_pos = cur_pos
        elif op == 3:                                           # DELETION
            edit = {"from_length": next_size}
        elif op == 2:                                           # MATCH
            edit = {"from_length": next_size, "to_length": next_size}
            qpos += next_size
        cigar_offset += next_size
        cur_pos = next_pos
        mapping["edit"].append(edit)
        if cigar_offset == n:
            it += 1
            cigar_offset = 0
    rank = 1
    mapping["rank"] = rank
    if not mapping["edit"]:
        del mapping["edit"]
    path["mapping"].append(mapping)
    for node in nodes[1:]:                                      # one character per further node
        rank += 1
        mapping = {"position": {"node_id": int(node), "offset": node_size - 1}, "edit": []}
        op, n = ops[it]
        if op in (4, 0):
            length = n - cigar_offset
            mapping["edit"].append({"to_length": length, "sequence": query_view[qpos:qpos + length]})
            qpos += length
            it += 1
            cigar_offset = 0
            op, n = ops[it]
        if op == 1:
            edit = {"from_length": 1, "to_length": 1, "sequence": query_view[qpos:qpos + 1]}
            qpos += 1
        elif op == 3:
            edit = {"from_length": 1}
        else:
            edit = {"from_length": 1, "to_length": 1}
            qpos += 1
        cigar_offset += 1
        if cigar_offset == n:
            cigar_offset = 0
            it += 1
        mapping["edit"].append(edit)
        mapping["rank"] = rank
        path["mapping"].append(mapping)
    path["length"] = len(nodes)
    if int(nodes[0]) == int(nodes[-1]):
        path["is_circular"] = True
    path["name"] = label
    return path


def alignment_to_json(a, strand_query, node_size, is_secondary, read_name, label=""):
    """Alignment::to_json (alignment.cpp:882-963). strand_query = the query strand `a` was aligned to
    (a None alignment gives the record of an unaligned read)."""
    out = {"name": read_name, "sequence": strand_query}
    if a is None:
        return out
    if a.sequence:
        out.setdefault("annotation", {})["ref_sequence"] = a.sequence
    if not a.query_len:
        return out
    out.setdefault("annotation", {})["cigar"] = a.get_cigar_string()
    view = strand_query[a.query_begin:a.query_begin + a.query_len]
    if len(a.nodes):
        out["path"] = _path_json(a.nodes, a.cigar, node_size, view, a.offset, label)
    out["score"] = int(a.score)
    if a.query_begin:
        out["query_position"] = int(a.query_begin)
        out["soft_clipped"] = True
    if is_secondary:
        out["is_secondary"] = True
    out["identity"] = a.get_num_matches() / a.query_len if a.query_len else 0.0
    out["read_mapped"] = bool(a.query_len)
    if a.orientation:
        out["read_on_reverse_strand"] = True
    return out


def format_alignment_json(header, paths, k):
    """cli/align.cpp:289-304, JSON branch: one line per alignment (without the trailing newline)"""
    lines = []
    for i, a in enumerate(paths):
        q = reverse_complement(paths.get_query()) if a.orientation else paths.get_query()
        lines.append(_json_value(alignment_to_json(a, q, k, i > 0, header)))
    if not paths:
        lines.append(_json_value(alignment_to_json(None, "", k, False, header)))
    return "\n".join(lines)


class AlignmentResults(list):
    """Alignments of one read (alignment.hpp:366-406)."""

    def __init__(self, query):
        super().__init__()
        # upper-cased copy, bytes >= 128 -> 127 (alignment.cpp:1357-1358)
        self.query = "".join(chr(127) if ord(c) >= 128 else c.upper() for c in query)

    def get_query(self):
        return self.query


def format_alignment(header, paths, min_path_score=0, with_nodes=False):
    """cli/align.cpp:254-307, TSV branch (without the trailing newline)"""
    out = "%s\t%s" % (header, paths.get_query())
    if not paths:
        return out + "\t*\t*\t%d\t*\t*\t*" % min_path_score
    for a in paths:
        out += "\t" + repr(a)
    if with_nodes:
        for a in paths:
            out += "\t" + ",".join(str(int(n)) for n in a.nodes)
    return out


def results_of_handle(L, res, seq_batch):
    """AlignmentResults of every read of a mgb_results_t (seq_batch: the (header, sequence) pairs it was made from)"""
    alns = L.mgb_results_alignments(res)
    out = []
    first, count = ctypes.c_uint64(), ctypes.c_uint32()
    for r, (header, seq) in enumerate(seq_batch):
        ar = AlignmentResults(seq if isinstance(seq, str) else seq.decode("latin1"))
        L.mgb_results_read_range(res, r, ctypes.byref(first), ctypes.byref(count))
        for i in range(first.value, first.value + count.value):
            x = alns[i]
            a = Alignment()
            a.orientation = bool(x.orientation); a.score = x.score; a.offset = x.offset
            a.query_begin = x.query_begin; a.query_len = x.query_len
            a.nodes = np.ctypeslib.as_array(x.nodes, shape=(x.num_nodes,)).copy() if (x.num_nodes and x.nodes) else np.zeros(0, np.uint64)
            a.sequence = ctypes.string_at(x.sequence, x.sequence_len).decode("latin1")
            a.cigar = np.ctypeslib.as_array(x.cigar, shape=(x.num_cigar_ops,)).copy()
            ar.append(a)
        out.append(ar)
    return out


class B200Aligner:
    """IDBGAligner over the B200 kernels (Seeder = SuffixSeeder<UniMEMSeeder>, Extender =
    DefaultColumnExtender, as DBGAligner<> in dbg_aligner.hpp:42-45)."""

    def __init__(self, graph, config=None):
        self.graph = graph
        self.config = config or DBGAlignerConfig()
        self._L = graph._L
        self.last_stats = None
        # the DBGAligner constructor throws for a configuration it cannot run (dbg_aligner.cpp:55-56)
        c = self.config.to_c() if not isinstance(self.config, mgb_config_t) else self.config
        _lib.check(self._L, self._L.mgb_config_check(self.graph.handle, ctypes.byref(c)))

    def get_graph(self):
        return self.graph

    def get_config(self):
        return self.config

    def has_coordinates(self):
        return False

    def align_batch_raw(self, buf, offsets):
        """Runs mgb_align_batch on packed reads; returns the opaque results handle."""
        c = self.config.to_c() if not isinstance(self.config, mgb_config_t) else self.config
        res = ctypes.c_void_p()
        _lib.check(self._L, self._L.mgb_align_batch(self.graph.handle, ctypes.byref(c), buf.ctypes.data,
                                                    offsets.ctypes.data, len(offsets) - 1, ctypes.byref(res)))
        return res

    def free_raw(self, res):
        self._L.mgb_results_free(res)

    def set_pipeline_pieces(self, max_pieces):
        """mgb_set_pipeline_pieces: cap on the pieces a batch is split into (0 = automatic)."""
        self._L.mgb_set_pipeline_pieces(int(max_pieces))

    def stats_of(self, res):
        st = self._L.mgb_results_stats(res).contents
        return {n: getattr(st, n) for n, _ in _lib.mgb_stats_t._fields_}

    def align_batch(self, seq_batch, callback=None):
        """seq_batch: list of (header, sequence). Calls callback(header, AlignmentResults) in input
        order (dbg_aligner.cpp:263, 353) and returns the list of AlignmentResults."""
        seqs = [s for _, s in seq_batch]
        buf, offsets = _pack(seqs)
        res = self.align_batch_raw(buf, offsets)
        try:
            self.last_stats = self.stats_of(res)
            out = results_of_handle(self._L, res, seq_batch)
            if callback:
                for (header, _), ar in zip(seq_batch, out):
                    callback(header, ar)
            return out
        finally:
            self.free_raw(res)

    def align(self, query):
        return self.align_batch([("", query)])[0]
