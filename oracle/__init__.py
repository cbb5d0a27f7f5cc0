"""CPU oracle for the DetectMate detector hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker (or as the timed CPU arm), never as
the thing shipped.  The product path (``detectmateservice_b200``) never imports
this package and fails loudly when its CUDA library is missing.

PARITY STATUS (see oracle/README.md and DESIGN.md):
  * wire schema (ParserSchema / DetectorSchema)      -- PINNED against the reference's
    embedded descriptor (container/fluentout/schemas_pb.rb:8) and the 201-byte
    fixture of tests/library_integration/library_integration_base_fixtures.py:27-43.
  * NewValueDetector output for the documented example  -- PINNED against the prose
    golden of docs/getting_started.md:423-435,498-510.
  * DummyDetector pattern                              -- PINNED
    (tests/library_integration/test_detector_integration.py:83-84,89-115).
  * NewValueDetector flags/scores in general           -- **PARITY UNPINNED**: the
    arithmetic lives in detectmatelibrary 0.1.0 @ ecdda558 (uv.lock:240-251) which
    is absent from /root/reference and not installable here; the reference's own
    tests assert nothing on its output (test_pipe_filereader_matcher_nvd.py:195-203).
    The restatement follows the documented contract (docs/interfaces.md:138-204,
    docs/getting_started.md:403-435).
"""
