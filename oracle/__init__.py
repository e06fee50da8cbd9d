"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Sin3DM triplane-denoising hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU baseline), never as the thing shipped.  The product path
(``sin3dm_b200``) never imports this package and fails loudly when its CUDA library is missing.

Parity pin: the restatement is checked against the *real* reference (imported read-only from
/root/reference/src in the authoring container by ``oracle/make_golden.py``); the resulting
vectors are committed under ``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.
The reference itself ships no tests / golden vectors, so these fixtures are the pin.
"""
