"""Minimal stand-in for `parameterized.parameterized_class` so the reference's own unittest files
run under pytest in the build container. Test infrastructure, not product."""
import sys


def parameterized_class(params):
    def decorator(base):
        module = sys.modules[base.__module__]
        testNames = [n for n in list(base.__dict__) if n.startswith("test")]
        for idx, p in enumerate(params):
            name = f"{base.__name__}_{idx}_{str(p.get('name', idx)).replace('-', '_')}"
            body = dict(p)
            for n in testNames:
                body[n] = base.__dict__[n]
            setattr(module, name, type(name, (base,), body))
        for n in testNames:
            delattr(base, n)
        return base

    return decorator
