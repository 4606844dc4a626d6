"""CPU oracle for the EWViT per-frame forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``efficient-wavelet-vit_b200/``) never imports it and
fails loudly when its CUDA library is missing.

PARITY UNPINNED at one boundary: the Haar analysis filter bank lives in the
third-party ``pytorch_wavelets`` package (un-pinned in the reference's
``requirements.txt:9``, not vendored, not installable here).  ``oracle/haar.py``
restates its published algorithm; the reference has no test or golden vector that
pins it.  Everything downstream of the DWT is checked
against the unmodified reference modules imported from ``/root/reference`` by
``tests/golden/make_golden.py``.
"""
