"""Regular package that shadows the reference's namespace package ``models`` (transformer_rawIQ/models/)."""
