"""TEST INFRASTRUCTURE ONLY -- stand-in for pytorch-ignite 0.5.0.post2 (reference requirements.txt:8).

Only control flow lives in ignite (no arithmetic): `Engine.run` iterates the loader once per epoch and fires
events; `Metric.attach` wires reset/update/compute.  Implemented from the documented behaviour; assumption
(SURVEY.md section 8c): the plain `Engine` never touches the global RNG and re-`iter()`s the loader each epoch.
"""
