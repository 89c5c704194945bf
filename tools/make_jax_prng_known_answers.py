#!/usr/bin/env python
"""Writes tests/golden/jax_prng_known_answers.json: PUBLIC known answers for the JAX PRNG restatement
(oracle/jax_prng.py, carl_b200/csrc/rng.h). JAX itself cannot be installed here, so -- as for gymnasium's CartPole
(tools/make_gymnasium_known_answers.py) -- the vectors are the published ones:

* Random123 (Salmon, Moraes, Dror, Shaw; SC'11) kat_vectors, generator `threefry2x32 20`: three (counter, key) ->
  output triples;
* the JAX documentation ("JAX - The Sharp Bits: random numbers" and the `jax.random` module docs), which print for
  `key = random.PRNGKey(0)`: `random.split(key)` -> [[4146024105, 967050713], [2718843009, 1272950319]],
  `random.normal(key, (1,))` -> [-0.20584226], `random.uniform(key)` -> 0.41845703, and for
  `key, subkey = random.split(key)`: `random.normal(subkey, (1,))` -> [-1.2515389].
"""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jax_prng_known_answers.json")
data = {
    "source": __doc__,
    "threefry2x32_20_kat": [
        {"counter": ["00000000", "00000000"], "key": ["00000000", "00000000"], "out": ["6b200159", "99ba4efe"]},
        {"counter": ["ffffffff", "ffffffff"], "key": ["ffffffff", "ffffffff"], "out": ["1cb996fc", "bb002be7"]},
        {"counter": ["243f6a88", "85a308d3"], "key": ["13198a2e", "03707344"], "out": ["c4923a9c", "483df7a0"]},
    ],
    "jax_docs_prngkey0": {
        "split": [[4146024105, 967050713], [2718843009, 1272950319]],
        "normal_1": -0.20584226,
        "uniform_scalar": 0.41845703,
        "subkey_normal_1": -1.2515389,
    },
}
json.dump(data, open(OUT, "w"), indent=1)
print(OUT)
