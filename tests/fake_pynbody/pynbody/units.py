import re


class Unit:
    def __init__(self, si, name="?"):
        self.si = float(si)
        self.name = name

    def __mul__(self, o):
        return Unit(self.si * o.si, f"({self.name} {o.name})")

    def __truediv__(self, o):
        return Unit(self.si / o.si, f"({self.name}/{o.name})")

    def __pow__(self, p):
        return Unit(self.si ** p, f"{self.name}**{p}")

    def ratio(self, other):
        return self.si / other.si

    def __repr__(self):
        return f"Unit({self.name})"


_BASE = {"m": 1.0, "km": 1.0e3, "kpc": 3.0856775814913673e19, "Mpc": 3.0856775814913673e22, "s": 1.0,
         "kg": 1.0, "Msol": 1.98847e30, "1": 1.0}
G = Unit(6.67430e-11, "G")  # m^3 kg^-1 s^-2
kpc = Unit(_BASE["kpc"], "kpc")
Msol = Unit(_BASE["Msol"], "Msol")
km = Unit(_BASE["km"], "km")


def parse(text):
    if isinstance(text, Unit):
        return text
    si = 1.0
    for tok in text.split():
        m = re.fullmatch(r"([A-Za-z0-9]+)(?:\*\*(-?\d+))?", tok)
        si *= _BASE[m.group(1)] ** int(m.group(2) or 1)
    return Unit(si, text)
