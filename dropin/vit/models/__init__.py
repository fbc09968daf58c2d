"""Regular package that shadows the reference's namespace package ``models`` (ViT/models/)."""
