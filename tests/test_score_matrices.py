"""tests/graph/test_aligner.cpp:29-76 (check_score_matrix_dna / _protein / _dna_unit / _protein_unit): every scoring matrix
the host mirror builds has positive match scores, the match as the row maximum, and is symmetric."""
import pytest

from metagraph_b200.config import blosum62_scoring_matrix, dna_scoring_matrix, unit_scoring_matrix

DNA5 = "ACGTN"                                   # kmer/alphabets.hpp kAlphabetDNA5
PROTEIN = "ABCDEFGHIJKLMNOPQRSTUVWYZX"           # kAlphabetProtein


@pytest.mark.parametrize("name,matrix,alphabet", [
    ("dna", dna_scoring_matrix(2, -1, -2), DNA5),
    ("protein", blosum62_scoring_matrix(), PROTEIN),
    ("dna_unit", unit_scoring_matrix(1, "ACGT"), DNA5),
    ("protein_unit", unit_scoring_matrix(1, "ABCDEFGHIJKLMNOPQRSTUVWYZ"), PROTEIN),
])
def test_check_score_matrix(name, matrix, alphabet):
    for i, a in enumerate(alphabet):
        last = i + 1 == len(alphabet)            # the wildcard letter is exempt from the match checks
        if not last:
            assert matrix[ord(a)][ord(a)] > 0
        for b in alphabet:
            if not last:
                assert matrix[ord(a)][ord(a)] >= matrix[ord(a)][ord(b)]
            assert matrix[ord(a)][ord(b)] == matrix[ord(b)][ord(a)]
