"""CPU oracle for the MC-parallel variational-layer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the CPU thing being timed.
The product path (``pytorch_probabilisticlayers_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the
pins are outputs of the reference itself, imported unmodified from /root/reference in the
builder container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function here against those fixtures.

Modules
  closed_form.py  numpy (fp64 by default) restatement of forward / backward / KL / losses
  philox.py       numpy restatement of the counter-based eps stream the kernels generate
  torch_port.py   op-for-op torch-CPU port of the reference layers + train step (cpu baseline)
"""
