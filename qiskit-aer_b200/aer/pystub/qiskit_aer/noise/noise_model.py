class NoiseModel:
    """Placeholder: noise models are passed as plain dicts (NoiseModel.to_dict() schema)."""

    def __init__(self, d=None):
        self._d = d or {"errors": []}

    def to_dict(self):
        return self._d
