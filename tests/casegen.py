"""Random (inserts, template) cases from a brace / escape / sentinel-heavy alphabet."""
import random

BS = "\\"
ALPHABET = ["{", "}", "{", "}", BS, ".", "a", "b", "-", " ", "〠", "n1", "k1", "k2", "k3", "miss", "é", "\n"]
KEYS = ["a", "b", "k1", "k2", "k3", "a-b", "b-a", "aa", "ab", "n1", "a.", ".a", "k1-v", "ARG1", "ARG", "HH:MM", "true", "7", "xy"]
VALUES = ["a", "b", "k1", "k2", "v", "", "a-b", "x y", "end.", "{a}", "{k1}", BS + "{q" + BS + "}", "." + BS + "}", "d." + BS + "}",
          BS + "}", "t" + BS, "}", "{", "a}b", "{k2}-{k3}", "〠.", ".〠", "{b}{a}", "{{k1}}", BS + "}" + BS + "}", "long value " * 9,
          "x" * 17, "é〠ü", 7, 0, -3, 12345678901234, True, False, None, ["x", "y"], ["a", 1, [True]], {"k": 1}, [], {}]


def gen_case(rng):
    ins = {}
    for k in rng.sample(KEYS, rng.randint(2, len(KEYS))):
        r = rng.random()
        if r < 0.8:
            ins[k] = rng.choice(VALUES)
        else:
            ins[k] = "".join(rng.choice(ALPHABET) for _ in range(rng.randint(0, 6)))
    mode = rng.random()
    if mode < 0.4:
        t = "".join(rng.choice(ALPHABET) for _ in range(rng.randint(0, 16)))
    else:
        parts = []
        for _ in range(rng.randint(1, 5)):
            k = rng.choice(KEYS + ["miss", "{k1}", "a-{k2}", "{k3}-b", "{{k1}}", "a{b}-{k1}", "", "{a}{b}"])
            parts.append(rng.choice(["", "x", ". ", BS + "{", BS + "}", "." + BS + "}", "lit ", "é"]) + "{" + k + "}")
        t = "".join(parts) + rng.choice(["", ".", BS + "}", "tail", "}", "{"])
        if mode > 0.9:
            t = "{" + t + "}"
    return ins, t


def gen_cases(seed, n):
    rng = random.Random(seed)
    return [gen_case(rng) for _ in range(n)]
