"""Alphabets of compact-genome 12.5.0 as used by tsalign (TEST INFRASTRUCTURE ONLY).

The reference takes them from the external crate `compact-genome`
(/root/reference/Cargo.toml:27); call sites: tsalign/src/align.rs:288-295,389-405,
lib_tsalign/src/a_star_aligner/alignment_result/a_star_sequences.rs:28-38.
Pinned by reference fixtures: index order A,C,G,T of DnaAlphabet
(lib_tsalign/src/costs/gap_affine/io/tests.rs:8-24) and the A<->T / C<->G
complement (reference_rc / query_rc of every test_files/*.toml).
PARITY UNPINNED: N self-complement, U for RNA and the IUPAC complement map are
restated from the IUPAC standard; no reference test exercises them.  Index order
beyond ACGT is unobservable on this path (tables are keyed by letter).
"""

ALPHABETS = {
    "dna": "ACGT",
    "dna-n": "ACGTN",
    "rna": "ACGU",
    "rna-n": "ACGUN",
    "dna-iupac": "ACGTRYSWKMBDHVN",
    "rna-iupac": "ACGURYSWKMBDHVN",
}

_COMPLEMENT = {
    "A": "T", "T": "A", "U": "A", "C": "G", "G": "C", "N": "N",
    "R": "Y", "Y": "R", "S": "S", "W": "W", "K": "M", "M": "K",
    "B": "V", "V": "B", "D": "H", "H": "D",
}


def chars(alphabet: str) -> str:
    return ALPHABETS[alphabet]


def complement_char(alphabet: str, c: str) -> str:
    r = _COMPLEMENT[c]
    if alphabet.startswith("rna") and r == "T":
        r = "U"
    return r


def complement_table(alphabet: str) -> list:
    cs = chars(alphabet)
    return [cs.index(complement_char(alphabet, c)) for c in cs]


def encode(alphabet: str, seq) -> bytes:
    """ASCII -> alphabet indices; raises ValueError on a non-alphabet character
    (VectorGenome::from_slice_u8, tsalign/src/align.rs:389-405)."""
    cs = chars(alphabet)
    if isinstance(seq, str):
        seq = seq.encode()
    lut = {ord(c): i for i, c in enumerate(cs)}
    try:
        return bytes(lut[b] for b in seq)
    except KeyError as e:
        raise ValueError(f"character {chr(e.args[0])!r} is not part of alphabet {alphabet}") from None


def reverse_complement(alphabet: str, seq: str) -> str:
    return "".join(complement_char(alphabet, c) for c in reversed(seq))
