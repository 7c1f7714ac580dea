"""CPU oracle for the OpenQuantumBase.jl master-equation RHS (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product path.  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  See ``oracle/oqb_oracle.py`` for the restatement itself.
"""
