"""Import stub: the reference pulls `phoneme_len` through a module that builds an espeak
backend at import time (data/phonemization/phonemes.py:5-6,59-65).  espeak is absent here
and the hot path never phonemizes, so this stub only has to import."""
