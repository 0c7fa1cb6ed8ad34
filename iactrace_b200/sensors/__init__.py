from .square import SquareSensor, DifferentiableSquareSensor
from .hexagonal import HexagonalSensor, DifferentiableHexagonalSensor

__all__ = ["SquareSensor", "HexagonalSensor", "DifferentiableSquareSensor", "DifferentiableHexagonalSensor"]
