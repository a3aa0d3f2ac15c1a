"""Exception types with the reference's names (exceptions.py:5 in the reference)."""


class QuantizationError(RuntimeError):
    """Raised for strict-quantization violations, re-initialisation errors and empty dynamic input."""


class ExportError(RuntimeError):
    pass
