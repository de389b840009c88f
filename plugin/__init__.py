"""WanGP model plugin: B200-native Wan denoise + VAE path (see plugin_info.json, docs/PLUGINS.md of the reference)."""
