"""Synthetic weights and inputs for the tests: re-exported from ``clipbert_b200.workload`` (the generators are shapes and value
ranges only; they live in the package so that the product arm of bench.py imports nothing from ``oracle/``)."""
from clipbert_b200.workload import (BERT_CFG, cnn_state_dict, full_state_dict, synth_batch, synth_images, synth_text,  # noqa: F401
                                    transformer_state_dict)
