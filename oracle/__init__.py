"""CPU oracle for the MSDA hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product package (``msda-triton_b200/msda_triton``) never does.
"""
