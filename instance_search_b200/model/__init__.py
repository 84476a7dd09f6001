"""Drop-in mirror of the reference's ``model/`` operator surface for the hot
path: ``custom_modules`` (NormalizeL2, Shift, TripletLoss), ``siamese``
(RegionDescriptorNet, DescriptorNet) and the net-surgery helpers they need."""
