"""CPU oracle for the Bulletproofs++ hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

A plain-Python (big-int) restatement of the reference algorithm
(Liam-Eagen/BulletproofsPP, pure Haskell), used only as the *checker* by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under
``bulletproofspp_b200/`` imports it.

PARITY UNPINNED.  The reference ships no golden vectors / known-answer tests
and cannot be built here (no GHC; elliptic-curve-0.3.0 / galois-field-1.0.1 are
not vendored).  Three transcript-affecting behaviours live in those packages
(``show`` of a field element, the root returned by ``pointX``, incomplete
projective addition); they are explicit policies in ``transcript.py``.  What
*is* pinned: SHA-256 (FIPS vectors), the secp256k1 group law (OpenSSL through
``cryptography``), the CM constants of FastSECP256K1.hs, the proof shapes of
README.md / the paper, and full prove->verify self-consistency on all eight
``examples/``.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
