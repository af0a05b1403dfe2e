class Parameter:
    def __init__(self, modelname, name, val, min=None, max=None, hard_min=None, hard_max=None,
                 units="", frozen=False, **kw):
        self.modelname, self.name, self.val, self.min, self.max = modelname, name, float(val), min, max
        self.units, self.frozen, self.fullname = units, frozen, f"{modelname}.{name}"


class RegriddableModel1D:
    def __init__(self, name, pars=()):
        self.name, self.pars = name, tuple(pars)

    def __call__(self, x):
        return self.calc([p.val for p in self.pars], x)
