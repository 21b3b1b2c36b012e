"""ORACLE — test infrastructure only.

CPU restatement of the TEOChat inference hot path, used ONLY by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker.  Nothing under ``teochat_b200/`` may import it.

How it is pinned.  The reference ships no tests, golden vectors or known-answer files for this path
 and its package cannot be imported whole here (it pins transformers==4.31.0 / peft /
decord, all absent).  Its own hot-path modules CAN be imported one by one with the unused third-party
imports mocked, so:
  * everything up to the language model — prompt, -200 tokenisation, torchvision transform chain,
    CLIPVisionTransformer, feature_select, mlp2x_gelu projector, prepare_inputs_labels_for_multimodal — is
    pinned against outputs of the REFERENCE'S OWN CODE run in this container
    (tests/golden/make_reference_golden.py → tests/golden/reference_path.*, checked by
    tests/test_reference_golden_cpu.py);
  * the LLaMA arithmetic lives in third-party transformers==4.31.0 (not vendored): PARITY UNPINNED against
    4.31 itself; the restatement follows its op order and is pinned against the
    installed transformers 5.5 ``LlamaForCausalLM(eager)`` / ``CLIPVisionModel`` on the same weights to
    1e-5 (tests/test_oracle_cpu.py) and through the reference-code fixture above (logits, greedy ids);
  * committed golden vectors of the oracle itself (tests/golden/make_golden.py) fix it over time.
"""
