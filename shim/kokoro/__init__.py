"""``kokoro`` import paths of igorshmukler/kokoro-ruslan, served by the B200 implementation (kokoro_ruslan_b200).

Only the surfaces of the accelerated hot path exist here: the ``kokoro-train`` console script,
``TrainingConfig``, ``KokoroModel``, the collate function and batch samplers, the length utilities and the HiFi-GAN
generator.  Everything else of the reference package (phoneme front-end, MFA, checkpoint manager, ...) is out of scope;
install this shim INSTEAD of the reference package, not next to it.
"""
__version__ = "0.0.35+b200"
