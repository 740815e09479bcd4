"""reference utils/decorator_peeling.py: reach the original function under a stack of ``functools.wraps`` decorators."""


def peel_decorations(decorated_function):
    fn = decorated_function
    while hasattr(fn, "__wrapped__"):
        fn = fn.__wrapped__
    return fn
