"""DBGAlignerConfig mirror (reference: metagraph/src/graph/alignment/aligner_config.hpp:18-94)
and its C-ABI POD image `mgb_config_t` (include/mgb.h).

Two factory functions give the two parameter sets the reference itself uses:
  * struct_defaults()  -- DBGAlignerConfig{} as the unit tests use it (aligner_config.hpp:23-54)
  * cli_defaults(k)    -- what `metagraph align` builds from its flags
                          (cli/config/config.hpp:114-145 + cli/align.cpp:33-69)
"""
import ctypes
from dataclasses import dataclass, field

INT32_MIN = -2**31
INT32_MAX = 2**31 - 1
SIZE_MAX = 2**64 - 1
DBL_MAX = 1.7976931348623157e308
NINF = INT32_MIN + 100


class mgb_config_t(ctypes.Structure):
    _fields_ = [
        ("num_alternative_paths", ctypes.c_uint64),
        ("min_seed_length", ctypes.c_uint64),
        ("max_seed_length", ctypes.c_uint64),
        ("max_num_seeds_per_locus", ctypes.c_uint64),
        ("min_cell_score", ctypes.c_int32),
        ("min_path_score", ctypes.c_int32),
        ("xdrop", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("min_exact_match", ctypes.c_double),
        ("max_nodes_per_seq_char", ctypes.c_double),
        ("max_ram_per_alignment", ctypes.c_double),
        ("rel_score_cutoff", ctypes.c_double),
        ("gap_opening_penalty", ctypes.c_int8),
        ("gap_extension_penalty", ctypes.c_int8),
        ("left_end_bonus", ctypes.c_int8),
        ("right_end_bonus", ctypes.c_int8),
        ("forward_and_reverse_complement", ctypes.c_uint8),
        ("global_xdrop", ctypes.c_uint8),
        ("allow_left_trim", ctypes.c_uint8),
        ("no_backtrack", ctypes.c_uint8),
        ("seed_complexity_filter", ctypes.c_uint8),
        ("result_nodes", ctypes.c_uint8),
        ("no_exact_path_shortcut", ctypes.c_uint8),
        ("reserved1", ctypes.c_uint8 * 5),
        ("score_matrix", (ctypes.c_int8 * 128) * 128),
    ]


def dna_scoring_matrix(match, transition, transversion):
    """aligner_config.cpp:164-183"""
    m = [[transversion] * 128 for _ in range(128)]
    for a, b in (("A", "G"), ("C", "T")):
        m[ord(a)][ord(b)] = transition
        m[ord(b)][ord(a)] = transition
    for c in "ACGT":
        m[ord(c)][ord(c)] = match
    return m


def unit_scoring_matrix(match, letters="ACGT"):
    """aligner_config.cpp:185-205 (valid upper-case letters only)"""
    m = [[-match] * 128 for _ in range(128)]
    for c in letters:
        m[ord(c)][ord(c)] = match
    return m


_B62_ORDER = "ARNDCQEGHILKMFPSTWYVBZX"
_B62 = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1
"""


def blosum62_scoring_matrix():
    """aligner_config.cpp:207-255: standard BLOSUM62, -4 elsewhere, +1 on the diagonal"""
    m = [[-4] * 128 for _ in range(128)]
    for i in range(128):
        m[i][i] = 1
    rows = [[int(x) for x in line.split()] for line in _B62.strip().splitlines()]
    for i, a in enumerate(_B62_ORDER):
        for j, b in enumerate(_B62_ORDER):
            m[ord(a)][ord(b)] = rows[i][j]
    return m


@dataclass
class DBGAlignerConfig:
    num_alternative_paths: int = 1
    min_seed_length: int = 0
    max_seed_length: int = 0
    max_num_seeds_per_locus: int = SIZE_MAX
    min_cell_score: int = NINF
    min_path_score: int = 0
    xdrop: int = INT32_MAX
    min_exact_match: float = 0.0
    max_nodes_per_seq_char: float = DBL_MAX
    max_ram_per_alignment: float = DBL_MAX
    rel_score_cutoff: float = 0.0
    gap_opening_penalty: int = -5
    gap_extension_penalty: int = -2
    left_end_bonus: int = 0
    right_end_bonus: int = 0
    forward_and_reverse_complement: bool = True
    global_xdrop: bool = True
    allow_left_trim: bool = True
    no_backtrack: bool = False
    # sdust is not vendored in the reference tree: the filter is restated from the symmetric-DUST definition
    # (parity with the library unpinned, tests/test_sdust.py); parity / bench runs keep it off
    seed_complexity_filter: bool = False
    # mgb_config_t::result_nodes: 0 = node paths come back with the alignments, 1 = they stay on the device
    result_nodes: int = 0
    # mgb_config_t::no_exact_path_shortcut: True = always run the extension (measurements; results are identical)
    no_exact_path_shortcut: bool = False
    score_matrix: list = field(default_factory=lambda: dna_scoring_matrix(2, -1, -2))

    def to_c(self):
        c = mgb_config_t()
        for name, _ in mgb_config_t._fields_:
            if name in ("reserved0", "reserved1", "score_matrix"):
                continue
            setattr(c, name, int(getattr(self, name)) if not isinstance(getattr(self, name), float)
                    else getattr(self, name))
        for i in range(128):
            for j in range(128):
                c.score_matrix[i][j] = self.score_matrix[i][j]
        return c


def struct_defaults(**kw):
    return DBGAlignerConfig(**kw)


def cli_defaults(k, alphabet="dna", edit_distance=False, **kw):
    """`metagraph align` defaults: match 2, mismatch -3/-3, gaps -6/-2, end bonus 5, xdrop 27,
    rel_score_cutoff 0.95, min_seed 19 (capped at k), max_seed inf, 1000 seeds/locus,
    5 nodes/char, 200 MB, min_exact_match 0.7; seed complexity filter off here (the CLI default is on; pass
    seed_complexity_filter=True for `metagraph align` without --align-no-seed-complexity-filter).
    alphabet="protein": BLOSUM62 (DBGAlignerConfig::set_scoring_matrix, aligner_config.cpp:164-205);
    the reverse-complement strand does not exist there (dbg_aligner.cpp:224-229)."""
    d = dict(num_alternative_paths=1, min_seed_length=min(19, k), max_seed_length=SIZE_MAX,
             max_num_seeds_per_locus=1000, min_path_score=0, xdrop=27, min_exact_match=0.7,
             max_nodes_per_seq_char=5.0, max_ram_per_alignment=200.0, rel_score_cutoff=0.95,
             gap_opening_penalty=-6, gap_extension_penalty=-2, left_end_bonus=5, right_end_bonus=5,
             forward_and_reverse_complement=True, seed_complexity_filter=False,
             score_matrix=dna_scoring_matrix(2, -3, -3))
    if alphabet == "protein":
        d["score_matrix"] = blosum62_scoring_matrix()
        d["forward_and_reverse_complement"] = False
    if edit_distance:        # --align-edit-distance (DBGAlignerConfig::set_scoring_matrix, aligner_config.cpp:128-148)
        d["score_matrix"] = unit_scoring_matrix(1, "ABCDEFGHIJKLMNOPQRSTUVWYZ" if alphabet == "protein" else "ACGT")
        d["left_end_bonus"] = d["right_end_bonus"] = 0
    d.update(kw)
    return DBGAlignerConfig(**d)
