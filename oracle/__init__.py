"""CPU oracle for the Wave-Mamba forward hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the reported CPU baseline.
The product package (``wave_mamba_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * everything except the selective scan is PINNED against the reference's own
    ``basicsr/archs/wavemamba_arch.py`` executed in the build container
    (``tools/make_golden.py`` -> ``tests/golden/*.npz``);
  * the selective scan itself is PARITY UNPINNED against ``mamba_ssm``'s CUDA kernel
    (third-party, unpinned version, absent from the container); it restates the
    package's published ``selective_scan_ref`` recurrence and is cross-checked against
    a float64 evaluation.
"""
