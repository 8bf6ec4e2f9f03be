-- Test script (ours, in the idiom of the reference's lua/fft.lua and lua/iq-tex-filtered.lua):
-- device -> nrf_iq_filter (low-pass, FIR on the GPU) -> nrf_fft -> texture.

function setup()
    device = nrf_device_new(97.5, "../rfdata/rf-200.500-big.raw")
    filter = nrf_iq_filter_new(device.sample_rate, 200e3, 51)
    fft = nrf_fft_new(1024, 8)
    shader = ngl_shader_new(GL_TRIANGLES, "", "")
    texture = ngl_texture_new(shader, "uTexture")
end

function draw()
    samples_buffer = nrf_device_get_samples_buffer(device)
    nrf_iq_filter_process(filter, samples_buffer)
    filter_buffer = nrf_iq_filter_get_buffer(filter)
    nrf_fft_process(fft, filter_buffer)
    fft_buffer = nrf_fft_get_buffer(fft)
    ngl_texture_update(texture, fft_buffer, 1024, 8)
end
