// placeholder until NRC oracle lands
