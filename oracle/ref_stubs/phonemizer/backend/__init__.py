class EspeakBackend:
    def __init__(self, *a, **k):
        pass

    def phonemize(self, texts, *a, **k):
        raise RuntimeError("espeak is not available in this environment")
