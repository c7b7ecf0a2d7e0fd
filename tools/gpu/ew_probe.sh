python tools/perf_probe.py relu exp gelu tanh bf16:relu bf16:exp bf16:gelu bf16:silu bf16:tanh bf16:sigmoid
