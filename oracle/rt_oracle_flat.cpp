// placeholder, filled below
