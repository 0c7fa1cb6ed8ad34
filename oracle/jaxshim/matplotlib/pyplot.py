def __getattr__(name):
    raise RuntimeError("matplotlib stub: display functions are not available")
