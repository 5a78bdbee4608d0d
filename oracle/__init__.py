"""CPU oracle of the ClipBERT hot path. TEST INFRASTRUCTURE ONLY — never imported by clipbert_b200/."""
