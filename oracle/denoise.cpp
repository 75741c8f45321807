// placeholder until SVGF oracle lands
