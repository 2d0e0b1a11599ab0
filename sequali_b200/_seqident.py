"""sequali_b200._seqident -- the reference's ``sequali._seqident`` (``_seqidentmodule.c``) over libsqgpu
(ctypes mirror; the CPython extension is ``sequali_b200/ext/_seqident.so``).

``sequence_identity(target, query, ...)`` is the reference call (one pair, one launch);
``sequence_identities(pairs, ...)`` sends any number of pairs to the device in one launch (one warp per pair),
which is how ``identify_sequence`` should be driven when a report annotates hundreds of overrepresented
sequences against the contaminant library."""
import ctypes as C
from typing import Iterable, List, Tuple

import numpy as np

from . import _lib


def _int8(x: int) -> int:
    """The reference hands the scores to its loops as int8_t (_seqidentmodule.c:338-340)."""
    x &= 0xff
    return x - 256 if x >= 128 else x


def _ascii(s, first):
    if not isinstance(s, str):
        raise TypeError(f"identify_sequence() argument {'1' if s is first else '2'} must be str, not {type(s).__name__}")
    b = s.encode("utf-8")
    if len(b) != len(s):
        raise ValueError(f"Only ascii strings are allowed. Got {first!r}")  # (the reference names the target in both cases)
    return b


def sequence_identities(pairs: Iterable[Tuple[str, str]], match_score=1, mismatch_penalty=-1,
                        deletion_penalty=-1, insertion_penalty=-1) -> List[float]:
    """Identity of every (target, query) pair: matched query letters of the best local alignment / len(query)."""
    targets, queries = [], []
    for target, query in pairs:
        t = _ascii(target, target)
        q = _ascii(query, target)
        if len(q) > 31:
            raise ValueError(f"Only query with lengths less than 32 are supported. Got {len(q)}")
        targets.append(t)
        queries.append(q)
    n = len(targets)
    if n == 0:
        return []
    t_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in targets], out=t_off[1:])
    q_off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum([len(q) for q in queries], out=q_off[1:])
    t_blob, q_blob = b"".join(targets), b"".join(queries)
    out = np.zeros(n, dtype=np.int32)
    ctx = _lib.Context.get()
    _lib.check(ctx.lib.sq_sequence_identity_batch(
        ctx.h, t_blob, t_off.ctypes.data, q_blob, q_off.ctypes.data, n, _int8(match_score), _int8(mismatch_penalty),
        _int8(deletion_penalty), _int8(insertion_penalty), out.ctypes.data), "sq_sequence_identity_batch")
    return [float(m) / len(q) if len(q) else float("nan") for m, q in zip(out.tolist(), queries)]


def sequence_identity(target: str, query: str, match_score=1, mismatch_penalty=-1, deletion_penalty=-1,
                      insertion_penalty=-1) -> float:
    """reference _seqidentmodule.c:279-343"""
    return sequence_identities([(target, query)], match_score, mismatch_penalty, deletion_penalty, insertion_penalty)[0]
