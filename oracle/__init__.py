"""CPU oracle for the CanonicalSg2Im scene-graph -> layout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``canonicalsg2im_b200/`` imports this
package; the only allowed callers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function is a restatement (numpy for the integer work, torch-CPU for the
floating point work) of one reference function and cites the reference
``file:line`` it follows.  The restatement is pinned against outputs of the
*unmodified* reference, executed in the authoring container by
``oracle/make_golden.py`` (which imports ``/root/reference``); the resulting
vectors are committed under ``tests/golden/`` so that the GPU box, where the
reference does not exist, can still check against them.

Parity status: PINNED for canonicalization (reference KAT vectors
``scripts/graphs_utils.py:159-174`` + reference runs), PINNED-BY-REFERENCE-RUN
for GraphTripleConv / Sg2LayoutModel / layout / crop (the reference has no
tests of its own for the floating point ops; the golden vectors are outputs of
the reference itself on seeded inputs).
"""

from . import canon, graph, layout  # noqa: F401
