"""CPU oracle for the GenPercept one-step hot path.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch fp32 restatement of ``GenPerceptPipeline.single_infer``
(/root/reference/genpercept/genpercept_pipeline.py:375-526) and of the third-party graphs it
drives (diffusers 0.26-0.29 ``AutoencoderKL`` / ``UNet2DConditionModel`` with the
stabilityai/stable-diffusion-2-1 configs, restated from SURVEY.md Appendix A because diffusers is
not installed and is not vendored under /root/reference) plus the in-tree DPT head
(/root/reference/genpercept/models/dpt_head.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product path (``genpercept_b200``) never does.

PARITY STATUS: the reference holds no tests or golden tensors for this path, and
its arithmetic lives in diffusers which cannot be imported here -> **parity unpinned** for the inside of the
diffusers blocks of the VAE/UNet graphs.  Pinned: the orchestration (single_infer / encode_rgb / decode_pred) against
the reference's OWN pipeline class, and the UNet top-level dataflow (skip stack, upsample_size, DPT taps) against the
reference's OWN CustomUNet2DConditionModel.forward, both executed here around this package's modules
(tests/test_oracle_vs_reference_pipeline.py), and (tests/test_oracle.py): the DPT head against the reference's own
class (imported from /root/reference through a 2-symbol diffusers shim; fixtures committed under
tests/golden/), the scheduler collapse against a literal restatement of ddim.py, parameter counts
against the published model sizes, the empty-text embedding fixture, ``blocks.Upsample2D`` bitwise against the copy
of diffusers' class the reference vendors (dpt_head.py:92-210), and ``imgproc`` against torchvision's own resize.
Independent cross-check (not the reference): the whole VAE encoder and decoder against the LDM / taming encoder and
decoder that HF transformers ships (ChameleonVQVAEEncoder, JanusVQVAEDecoder) with the oracle's weights under the LDM
names — 4e-6 relative (tests/test_oracle_vs_ldm_vae.py).  Still restated only from SURVEY.md App. A: the inside of
the UNet's ResnetBlock2D (time-embedding add) and Transformer2DModel / BasicTransformerBlock.
"""
