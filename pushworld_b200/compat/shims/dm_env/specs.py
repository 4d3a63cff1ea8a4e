import numpy as np


class Array:
    def __init__(self, shape, dtype, name=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.name = name

    def validate(self, value):
        value = np.asarray(value)
        if value.shape != self.shape:
            raise ValueError(f"Expected shape {self.shape} but found {value.shape}")
        if value.dtype != self.dtype:
            raise ValueError(f"Expected dtype {self.dtype} but found {value.dtype}")
        return value


class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum = np.asarray(minimum, dtype=self.dtype)
        self.maximum = np.asarray(maximum, dtype=self.dtype)

    def validate(self, value):
        value = super().validate(value)
        if (value < self.minimum).any() or (value > self.maximum).any():
            raise ValueError("Values were not all within bounds")
        return value


class DiscreteArray(BoundedArray):
    def __init__(self, num_values, dtype=np.int32, name=None):
        super().__init__((), dtype, 0, num_values - 1, name)
        self.num_values = num_values
