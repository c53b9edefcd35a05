class _Style:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Font(_Style): pass
class Color(_Style): pass
class PatternFill(_Style): pass
class GradientFill(_Style): pass
class Border(_Style): pass
class Side(_Style): pass
class Alignment(_Style): pass
class Protection(_Style): pass
