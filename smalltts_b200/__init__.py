"""smalltts_b200: B200-native engine for the smalltts synthesize hot path."""


def __getattr__(name):  # lazy, like the reference's smalltts/__init__.py:1-5
    if name == "SmallTTS":
        from .infer import SmallTTS

        return SmallTTS
    if name == "Engine":
        from .engine import Engine

        return Engine
    raise AttributeError(name)
