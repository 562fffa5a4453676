"""TEST INFRASTRUCTURE ONLY -- empty stand-in for the reference's root `utils.py`.

`/root/reference/models/ode_transformer_gpt.py:3` does `import utils` and never uses it; the
real module needs torch_pca / matplotlib / imageio, none of which exist in this image."""
