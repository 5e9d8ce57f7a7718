from .time_generator import TimeGenerator, DefaultTimeGenerator, NormalTimeGenerator
