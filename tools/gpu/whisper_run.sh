cd axiom_b200/host/_build; ./whisper_encoder 8 5 1 | tail -1; ./whisper_encoder 32 3 0 | tail -1
