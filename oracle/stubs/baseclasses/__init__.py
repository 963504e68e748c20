"""Minimal stand-in for mdolab-baseclasses, used ONLY to import the reference FEMpy
package inside the build container (oracle/ref_harness.py). Test infrastructure, not product.

FEMpy needs `BaseSolver(name, category, defaultOptions=..., options=...)` with
`getOption/setOption/pp` (reference FEMpy/Model.py:82, FEMpy/Problem.py:89).
Default options are `[type, value]` pairs (reference FEMpy/Model.py:419-427).
"""


class BaseSolver:
    def __init__(self, name, category, defaultOptions=None, options=None, **kwargs):
        self.name = name
        self.category = category
        self.options = {}
        for key, spec in (defaultOptions or {}).items():
            self.options[key.lower()] = spec[1]
        for key, value in (options or {}).items():
            if key.lower() not in self.options:
                raise KeyError(f"Option {key} is not a valid {name} option")
            self.options[key.lower()] = value

    def getOption(self, name):
        return self.options[name.lower()]

    def setOption(self, name, value):
        self.options[name.lower()] = value

    def pp(self, obj):
        pass
