"""Undo ``functools.wraps`` decoration chains (role of the reference's utils/decorator_peeling.py)."""
import inspect


def peel_decorations(decorated_function):
    """The innermost callable under any number of ``functools.wraps``-style wrappers (follows ``__wrapped__`` links)."""
    return inspect.unwrap(decorated_function)
