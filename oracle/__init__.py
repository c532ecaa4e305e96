"""Parity oracle for the NAFAE grounding hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg and ``__graft_entry__.smoke()`` may import it;
``nafae_b200`` never does (tests/test_boundary.py greps for that).

* ``oracle.cpu``      ctypes wrapper of ``libnafae_oracle.so`` (nafae_oracle.c): NMS, proposal
                      tail, RoIAlign(+avg/max), RoIPool, forward + backward, on the host.
* ``oracle.dvsa``     torch-fp32 CPU restatement of ``DVSA.forward`` (reference model.py:517-614)
                      and ``postprocess`` / ``record_det`` (model.py:457-487).
* ``oracle.ref_gpu``  ctypes wrapper of ``oracle/_ref/libnafae_ref.so``: the reference's own,
                      unmodified ``.cu`` files compiled for sm_100a (GPU box only).
"""
