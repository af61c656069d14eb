"""RESULTS codes in definition order (reference lineax/_solution.py:52-68)."""


class RESULTS:
    successful = 0
    max_steps_reached = 1
    singular = 2
    breakdown = 3
    stagnation = 4
    conlim = 5
    nonfinite_input = 6

    names = [
        "successful", "max_steps_reached", "singular", "breakdown",
        "stagnation", "conlim", "nonfinite_input",
    ]
