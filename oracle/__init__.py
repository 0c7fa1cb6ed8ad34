"""CPU oracle for the IACTrace Monte-Carlo ray-tracing hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU with NumPy, the
algorithm of the reference (GerritRo/iactrace v0.4.0, ``/root/reference``) for
the path named in ``BASELINE.json``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the
reported CPU baseline -- the product package ``iactrace_b200`` never does.

Pinning status
--------------
* The reference ships no tests, golden vectors or stored arrays (SURVEY.md
  section 4), and JAX/Equinox are not installable in this image.
* ``oracle/prng.py`` is pinned on public JAX known-answer values
  (``tests/test_oracle_prng.py``).
* ``oracle/sample.py`` and ``oracle/trace.py`` are pinned on fixtures produced
  by executing the reference's own, unmodified Python sources from
  ``/root/reference`` on top of ``oracle/jaxshim`` (a NumPy stand-in for the
  small JAX/Equinox surface the reference uses); see
  ``tests/golden/make_golden.py``.  Because the arithmetic of real
  JAX/XLA is still substituted by NumPy, DESIGN.md records this as
  "reference logic pinned, XLA arithmetic unpinned".
"""
