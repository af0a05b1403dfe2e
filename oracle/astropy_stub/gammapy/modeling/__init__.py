import astropy.units as u


class Parameter:
    def __init__(self, name, value, unit="", min=None, max=None, frozen=False, **kw):
        self.name, self.value, self.unit, self.min, self.max, self.frozen = name, value, unit, min, max, frozen

    @property
    def quantity(self):
        return self.value * u.Unit(self.unit)


class Parameters(list):
    def __init__(self, parameters=()):
        super().__init__(parameters)

    @property
    def names(self):
        return [p.name for p in self]

    def __getitem__(self, key):
        if isinstance(key, str):
            for p in self:
                if p.name == key:
                    return p
            raise KeyError(key)
        return list.__getitem__(self, key)

    def select(self, names):
        return Parameters([p for p in self if p.name in names])
