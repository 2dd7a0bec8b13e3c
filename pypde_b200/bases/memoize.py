"""Caching decorator with the reference's name and behaviour (pypde/bases/memoize.py:10-38):
results are cached per argument tuple (for methods: per instance); unhashable
arguments bypass the cache."""
import functools


class memoized(object):
    def __init__(self, func):
        self.func = func
        self.cache = {}
        functools.update_wrapper(self, func)

    def __call__(self, *args):
        try:
            hash(args)
        except TypeError:
            return self.func(*args)
        if args not in self.cache:
            self.cache[args] = self.func(*args)
        return self.cache[args]

    def __get__(self, obj, objtype=None):
        return functools.partial(self.__call__, obj)
